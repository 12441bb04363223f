"""Fabricate the offline assets a recipe run needs: an HF-style llm_path (config.json + tokenizer), 16 kHz wavs, jsonl."""
import json
import os
import wave

import numpy as np


def make_llm_dir(path, vocab=512, hidden=256, layers=2, heads=4, kv_heads=2, ffn=512):
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    os.makedirs(path, exist_ok=True)
    words = ["<unk>", "<s>", "</s>"] + [chr(c) for c in range(32, 127)] + [f"w{i}" for i in range(vocab - 98)]
    tok = Tokenizer(models.WordLevel({w: i for i, w in enumerate(words)}, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.Split("", "isolated")
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="<unk>", bos_token="<s>", eos_token="</s>")
    fast.save_pretrained(path)
    cfg = dict(model_type="llama", architectures=["LlamaForCausalLM"], vocab_size=vocab, hidden_size=hidden, intermediate_size=ffn,
               num_hidden_layers=layers, num_attention_heads=heads, num_key_value_heads=kv_heads, rms_norm_eps=1e-5, rope_theta=10000.0,
               max_position_embeddings=4096, tie_word_embeddings=False)
    json.dump(cfg, open(os.path.join(path, "config.json"), "w"))
    return path


def make_data(path, n=6, seed=0):
    os.makedirs(path, exist_ok=True)
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(n):
        sec = float(rng.uniform(0.5, 2.0))
        pcm = (rng.standard_normal(int(16000 * sec)) * 3000).astype(np.int16)
        wav = os.path.join(path, f"utt{i}.wav")
        with wave.open(wav, "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
        rows.append({"key": f"utt{i}", "source": wav, "target": "hello world " + "ab" * (i % 3)})
    jl = os.path.join(path, "data.jsonl")
    with open(jl, "w") as f:
        f.write("\n".join(json.dumps(r) for r in rows))
    return jl


def run_config(llm_dir, jsonl, out_dir, model_file, dataset_file, **train_over):
    """The config tree of examples/asr_librispeech/asr_config.py (the keys the hot path reads)."""
    from omegaconf import OmegaConf
    train = dict(model_name="asr", enable_ddp=False, enable_deepspeed=False, enable_fsdp=False, low_cpu_fsdp=False, run_validation=True,
                 batch_size_training=2, batching_strategy="custom", context_length=4096, gradient_accumulation_steps=1, num_epochs=1,
                 num_workers_dataloader=0, warmup_steps=2, total_steps=100, validation_interval=2, lr=1e-3, weight_decay=0.0, seed=42,
                 use_fp16=False, mixed_precision=True, val_batch_size=2, use_peft=True,
                 peft_config=dict(peft_method="lora", r=8, lora_alpha=32, target_modules=["q_proj", "v_proj"], bias="none",
                                  task_type="CAUSAL_LM", lora_dropout=0.0, inference_mode=False),
                 output_dir=out_dir, freeze_layers=False, num_freeze_layers=1, quantization=False, one_gpu=False, save_model=True,
                 save_optimizer=False, use_fast_kernels=False, run_test_during_validation=False, freeze_llm=True, freeze_encoder=True)
    train.update(train_over)
    cfg = dict(
        dataset_config=dict(dataset="speech_dataset", file=dataset_file, train_data_path=jsonl, val_data_path=jsonl, prompt="Transcribe speech to text. ",
                            fix_length_audio=-1, inference_mode=False, input_type="mel", mel_size=80, normalize=False),
        model_config=dict(file=model_file, llm_name="tiny-llama-test", llm_path=llm_dir, llm_type="decoder_only", llm_dim=256, encoder_name="whisper",
                          encoder_ds_rate=2, encoder_path="tiny", encoder_dim=384, encoder_projector="linear", encoder_projector_ds_rate=5,
                          modal="audio", normalize=False, encoder_type="finetune", b200_random_init=True),
        train_config=train,
        log_config=dict(use_wandb=False, wandb_dir=out_dir, wandb_entity_name="x", wandb_project_name="x", wandb_exp_name="x",
                        log_file=os.path.join(out_dir, "train.log"), log_interval=5),
        fsdp_config=dict(mixed_precision=True, use_fp16=False, sharding_strategy="NO_SHARD", checkpoint_type="SHARDED_STATE_DICT",
                         fsdp_activation_checkpointing=True, fsdp_cpu_offload=False, pure_bf16=False, optimizer="AdamW"),
        debug=False, metric="acc", ckpt_path=None)
    return OmegaConf.create(cfg)
