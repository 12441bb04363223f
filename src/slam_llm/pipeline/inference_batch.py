"""Batch decoding entry point (reference: src/slam_llm/pipeline/inference_batch.py:22-135): same config split and seeding, plugin-loaded
model factory and dataset (split "test", inference-mode collator), `model.generate(**batch)` per batch, `tokenizer.batch_decode`, and the
`<decode_log>_pred` / `<decode_log>_gt` files (key \\t text).  The decode runs on the B200 decoder through slam_model.generate."""
import logging
import os
import random

import hydra
import torch
from omegaconf import DictConfig, OmegaConf
from tqdm import tqdm

from slam_llm.utils.dataset_utils import get_preprocessed_dataset
from slam_llm.utils.model_utils import get_custom_model_factory


@hydra.main(config_name=None, version_base=None)
def main_hydra(cfg: DictConfig):
    kwargs = cfg
    logging.basicConfig(level=getattr(logging, kwargs.get("log_level", "INFO").upper()))
    main(kwargs)


def main(kwargs: DictConfig):
    train_config, fsdp_config, model_config, log_config, dataset_config = (kwargs.train_config, kwargs.fsdp_config, kwargs.model_config,
                                                                          kwargs.log_config, kwargs.dataset_config)
    OmegaConf.set_struct(kwargs, False)
    for k in ("train_config", "fsdp_config", "model_config", "log_config", "dataset_config"):
        del kwargs[k]
    OmegaConf.set_struct(kwargs, True)
    log_dir = os.path.dirname(log_config.log_file)
    if log_dir and not os.path.exists(log_dir):
        os.makedirs(log_dir, exist_ok=True)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s | %(levelname)s | %(name)s | %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
    logger = logging.getLogger()
    logger.setLevel(logging.INFO)
    file_handler = logging.FileHandler(filename=log_config.log_file, mode="w")
    file_handler.setFormatter(logging.Formatter("[%(asctime)s][%(name)s][%(levelname)s] - %(message)s", datefmt="%Y-%m-%d %H:%M:%S"))
    logger.addHandler(file_handler)
    logger.info("train_config: {}".format(train_config))
    logger.info("model_config: {}".format(model_config))

    if not torch.cuda.is_available():
        raise RuntimeError("slam_llm.pipeline.inference_batch: no CUDA device; the B200 decoder has no CPU fallback")
    torch.cuda.manual_seed(train_config.seed)
    torch.manual_seed(train_config.seed)
    random.seed(train_config.seed)

    model_factory = get_custom_model_factory(model_config, logger)
    model, tokenizer = model_factory(train_config, model_config, **kwargs)
    device = torch.device("cuda")
    model.to(device)                                                       # (no-op: the step is built on the GPU)
    model.eval()

    logger.info("dataset_config: {}".format(dataset_config))
    dataset_test = get_preprocessed_dataset(tokenizer, dataset_config, split="test")
    dynamic = train_config.batching_strategy == "dynamic"
    if not dynamic:
        logger.info(f"--> Training Set Length = {len(dataset_test)}")
    test_dataloader = torch.utils.data.DataLoader(dataset_test, num_workers=train_config.num_workers_dataloader, pin_memory=True, shuffle=False,
                                                  batch_size=train_config.val_batch_size, drop_last=False, collate_fn=dataset_test.collator)
    logger.info("=====================================")
    pred_path = kwargs.get("decode_log") + "_pred"
    gt_path = kwargs.get("decode_log") + "_gt"
    with open(pred_path, "w") as pred, open(gt_path, "w") as gt:
        for step, batch in tqdm(enumerate(test_dataloader), total=len(test_dataloader) if not dynamic else None):
            for key in batch.keys():
                batch[key] = batch[key].to(device) if isinstance(batch[key], torch.Tensor) else batch[key]
            model_outputs = model.generate(**batch)
            output_text = model.tokenizer.batch_decode(model_outputs, add_special_tokens=False, skip_special_tokens=True)
            for key, text, target in zip(batch["keys"], output_text, batch["targets"]):
                pred.write(key + "\t" + text.replace("\n", " ") + "\n")
                gt.write(key + "\t" + target + "\n")
    return pred_path, gt_path


if __name__ == "__main__":
    main_hydra()
