"""Peak-memory tracker used around every epoch (reference: src/slam_llm/utils/memory_utils.py:13-61; same attribute names, in GiB),
device-agnostic: without CUDA only the host numbers are filled (reference quirk Q3)."""
import gc

import psutil
import torch

_GIB = 2 ** 30


def byte2gb(x):
    return int(x / _GIB)


def _host_rss_gb() -> int:
    return byte2gb(psutil.Process().memory_info().rss)


class MemoryTrace:
    """with MemoryTrace() as t: ...  ->  t.peak, t.max_reserved, t.peak_active_gb, t.cuda_malloc_retires, t.cpu_peaked (+ begin values)."""

    def __enter__(self):
        gc.collect()
        self.cuda = torch.cuda.is_available()
        self.begin = 0
        if self.cuda:
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
            self.begin = byte2gb(torch.cuda.memory_allocated())
        self.cpu_begin = _host_rss_gb()
        return self

    def __exit__(self, *exc):
        gc.collect()
        device = dict(peak=0, max_reserved=0, peak_active_gb=0, cuda_malloc_retires=0)
        if self.cuda:
            torch.cuda.empty_cache()
            stats = torch.cuda.memory_stats()
            device = dict(peak=byte2gb(torch.cuda.max_memory_allocated()), max_reserved=byte2gb(torch.cuda.max_memory_reserved()),
                          peak_active_gb=byte2gb(stats.get("active_bytes.all.peak", 0)), cuda_malloc_retires=stats.get("num_alloc_retries", 0))
        for name, value in device.items():
            setattr(self, name, value)
        self.cpu_end = _host_rss_gb()
        self.cpu_peaked = max(0, self.cpu_end - self.cpu_begin)
