"""String-addressed dataset plugins: `dataset_config.file = "path/to/module.py[:factory]"`
(reference behaviour: src/slam_llm/utils/dataset_utils.py:14-57 - same entry points, defaults and exception types)."""
import importlib.util
import logging
import os

logger = logging.getLogger(__name__)
DEFAULT_FACTORY = "get_custom_dataset"


def load_module_from_py_file(py_file: str) -> object:
    """Execute a .py file that is not on sys.path and return it as a module object."""
    spec = importlib.util.spec_from_file_location(os.path.basename(py_file), py_file)
    if spec is None or spec.loader is None:
        raise ImportError(f"cannot import {py_file}")
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def _plugin_factory(address: str, default: str = DEFAULT_FACTORY, kind: str = "dataset"):
    """"file.py:func" (or "file.py" -> `default`) -> the callable; shared by the dataset and the model plugin loaders."""
    file_part, _, func_part = address.partition(":")
    func_part = func_part or default
    if os.path.splitext(file_part)[1] != ".py":
        raise ValueError(f"Dataset file {file_part} is not a .py file.")
    if not os.path.isfile(file_part):
        raise FileNotFoundError(f"Dataset py file {file_part} does not exist or is not a file.")
    namespace = load_module_from_py_file(file_part)
    if not hasattr(namespace, func_part):
        logger.info(f"It seems like the given method name ({func_part}) is not present in the {kind} .py file ({file_part}).")
        raise AttributeError(f"module {file_part!r} has no attribute {func_part!r}")
    return getattr(namespace, func_part)


def get_custom_dataset(dataset_config, tokenizer, split: str):
    return _plugin_factory(dataset_config.file)(dataset_config, tokenizer, split)


def get_preprocessed_dataset(tokenizer, dataset_config, split: str = "train"):
    """The recipes only ever configure custom datasets; the argument order differs from get_custom_dataset as in the reference."""
    return get_custom_dataset(dataset_config, tokenizer, split)
