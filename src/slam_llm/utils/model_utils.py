"""String-addressed model plugins: `model_config.file = "path/to/model.py:model_factory"` (reference: src/slam_llm/utils/model_utils.py:4-29)."""
from slam_llm.utils.dataset_utils import _plugin_factory


def get_custom_model_factory(model_config, logger):
    address = model_config.get("file", None)
    if address is None:                                    # no plugin configured: the stock ASR model
        from slam_llm.models.slam_model import model_factory
        return model_factory
    return _plugin_factory(address, default="model_factory", kind="model")
