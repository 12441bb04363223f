"""String-addressed dataset plugins (reference: src/slam_llm/utils/dataset_utils.py:14-57)."""
import importlib.machinery
import importlib.util
import logging
from pathlib import Path

logger = logging.getLogger(__name__)


def load_module_from_py_file(py_file: str) -> object:
    """Import a module from a .py path that is not on sys.path."""
    name = Path(py_file).name
    loader = importlib.machinery.SourceFileLoader(name, py_file)
    spec = importlib.util.spec_from_loader(name, loader)
    module = importlib.util.module_from_spec(spec)
    loader.exec_module(module)
    return module


def _resolve(spec: str, default_func: str, kind: str):
    module_path, func_name = spec.split(":") if ":" in spec else (spec, default_func)
    if not module_path.endswith(".py"):
        raise ValueError(f"Dataset file {module_path} is not a .py file.")
    path = Path(module_path)
    if not path.is_file():
        raise FileNotFoundError(f"Dataset py file {path.as_posix()} does not exist or is not a file.")
    module = load_module_from_py_file(path.as_posix())
    try:
        return getattr(module, func_name)
    except AttributeError:
        logger.info(f"It seems like the given method name ({func_name}) is not present in the {kind} .py file ({path.as_posix()}).")
        raise


def get_custom_dataset(dataset_config, tokenizer, split: str):
    return _resolve(dataset_config.file, "get_custom_dataset", "dataset")(dataset_config, tokenizer, split)


def get_preprocessed_dataset(tokenizer, dataset_config, split: str = "train"):
    return get_custom_dataset(dataset_config, tokenizer, split)
