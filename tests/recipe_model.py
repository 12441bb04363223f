"""A recipe-style model plugin for the tests (same shape as the reference's examples/asr_librispeech/model/slam_model_asr.py:
a model_factory built from the slam_llm setup_* functions and a slam_model subclass)."""
import os

import torch

from slam_llm.models.slam_model import setup_encoder, setup_encoder_projector, setup_llm, setup_tokenizer, slam_model
from slam_llm.utils.train_utils import print_model_size


class slam_model_test_asr(slam_model):
    def __init__(self, encoder, llm, encoder_projector, tokenizer, train_config, model_config, **kwargs):
        super().__init__(encoder, llm, encoder_projector, tokenizer, train_config, model_config, **kwargs)


def model_factory(train_config, model_config, **kwargs):
    tokenizer = setup_tokenizer(train_config, model_config, **kwargs)
    encoder = setup_encoder(train_config, model_config, **kwargs)
    llm = setup_llm(train_config, model_config, **kwargs)
    encoder_projector = setup_encoder_projector(train_config, model_config, **kwargs)
    model = slam_model_test_asr(encoder, llm, encoder_projector, tokenizer, train_config, model_config, **kwargs)
    ckpt_path = kwargs.get("ckpt_path", None)
    if ckpt_path is not None:
        model.load_state_dict(torch.load(ckpt_path, map_location="cpu"), strict=False)
    print_model_size(model, train_config, int(os.environ["RANK"]) if train_config.enable_fsdp or train_config.enable_ddp else 0)
    return model, tokenizer
