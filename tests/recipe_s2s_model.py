"""A recipe-style s2s model plugin for the tests: the forward of a SLAM-Omni recipe (examples/s2s/model/slam_model_s2s.py:160-306 — multi-layer
token input averaged over `code_layer + 1` embedding lookups, audio features merged into the audio layers, logits over the expanded vocabulary
split into a text slice and per-layer audio slices, group cross-entropy) written against the slam_llm surface the reference recipe uses:
`self.encoder.extract_variable_length_features`, `self.encoder_projector(...)`, `self.llm.model.embed_tokens(...)`, `self.llm(inputs_embeds=...,
attention_mask=..., labels=...)` -> `.logits`, `compute_accuracy`.  Test infrastructure (the reference recipe itself needs snac / its own utils)."""
import torch
import torch.nn.functional as F

from slam_llm.models.slam_model import compute_accuracy, setup_encoder, setup_encoder_projector, setup_llm, slam_model


class slam_model_s2s_test(slam_model):
    def __init__(self, encoder, llm, encoder_projector, tokenizer, train_config, model_config, **kwargs):
        super().__init__(encoder, llm, encoder_projector, tokenizer, train_config, model_config, **kwargs)
        vc = model_config.vocab_config
        self.code_layer, self.text_vocab, self.audio_vocab = vc.code_layer, vc.padded_text_vocabsize, vc.padded_audio_vocabsize
        if vc.total_vocabsize != self.llm.lm_head.weight.size(0):
            self.llm.resize_token_embeddings(vc.total_vocabsize)
        self.group_decode_adapter = None
        if model_config.get("group_decode", False):          # a plain torch module OUTSIDE the arena: trained by FlatAdamW's inner AdamW
            n = self.code_layer * self.audio_vocab
            self.group_decode_adapter = torch.nn.Linear(n, n, bias=False, device="cuda")

    def forward(self, input_ids=None, attention_mask=None, labels=None, **kwargs):
        audio_pcm, modality_mask = kwargs.get("audio_pcm"), kwargs.get("modality_mask")
        L = self.code_layer
        mel = self.b200.log_mel(audio_pcm.to(self.b200.device, torch.float32))
        encoder_outs = self.encoder_projector(self.encoder.extract_variable_length_features(mel.permute(0, 2, 1))).float()
        input_ids = input_ids.clone()
        input_ids[input_ids == -1] = 0
        embed = self.llm.model.embed_tokens if hasattr(self.llm.model, "embed_tokens") else self.llm.model.model.embed_tokens
        inputs_embeds = embed(input_ids).float()                                      # [B, L + 1, S, D]
        mm = modality_mask.unsqueeze(1).repeat(1, L, 1)
        starts = (mm == True).float().argmax(dim=2)                                   # noqa: E712
        lengths = torch.clamp(mm.sum(dim=2), max=encoder_outs.shape[1]).tolist()
        pad = torch.zeros_like(inputs_embeds)
        for i in range(encoder_outs.shape[0]):
            for j in range(L):
                s0, n = starts[i, j].item(), lengths[i][j]
                pad[i, j, s0:s0 + n] = encoder_outs[i, :n]
        audio_layers = pad[:, :L] + inputs_embeds[:, :L] * (~mm[:, :, :, None])
        inputs_embeds = torch.cat([audio_layers, inputs_embeds[:, L:]], dim=1).mean(dim=1)
        text_labels, audio_labels = labels[:, L], labels[:, :L]
        model_outputs = self.llm(inputs_embeds=inputs_embeds, attention_mask=attention_mask, labels=text_labels)
        x = model_outputs.logits
        xt = x[..., : self.text_vocab]
        x_audio = x[..., self.text_vocab:]
        if self.group_decode_adapter is not None:
            x_audio = self.group_decode_adapter(x_audio)
        xa = [x_audio[..., i * self.audio_vocab:(i + 1) * self.audio_vocab] for i in range(L)]
        losses = [F.cross_entropy(xa[i][:, :-1].reshape(-1, self.audio_vocab), audio_labels[:, i, 1:].reshape(-1), ignore_index=-100) for i in range(L)]
        losses.append(F.cross_entropy(xt[:, :-1].reshape(-1, self.text_vocab), text_labels[:, 1:].reshape(-1), ignore_index=-100))
        model_outputs.loss = sum(losses) / (L + 1)
        with torch.no_grad():
            text_acc = compute_accuracy(torch.argmax(xt, -1)[:, :-1], text_labels[:, 1:], ignore_label=-100)
        return model_outputs, text_acc, [-1] * L, losses


def model_factory(train_config, model_config, **kwargs):
    encoder = setup_encoder(train_config, model_config, **kwargs)
    llm = setup_llm(train_config, model_config, **kwargs)
    projector = setup_encoder_projector(train_config, model_config, **kwargs)
    model = slam_model_s2s_test(encoder, llm, projector, None, train_config, model_config, **kwargs)
    ckpt_path = kwargs.get("ckpt_path", None)
    if ckpt_path is not None:
        model.load_state_dict(torch.load(ckpt_path, map_location="cpu"), strict=False)
    return model, None
