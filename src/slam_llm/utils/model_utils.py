"""String-addressed model plugins (reference: src/slam_llm/utils/model_utils.py:4-29)."""
from slam_llm.utils.dataset_utils import _resolve


def get_custom_model_factory(model_config, logger):
    path = model_config.get("file", None)
    if path is None:
        from slam_llm.models.slam_model import model_factory
        return model_factory
    return _resolve(path, "model_factory", "model")
