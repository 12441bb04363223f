"""Peak-memory tracker (reference: src/slam_llm/utils/memory_utils.py:13-61), device-agnostic (reference quirk Q3)."""
import gc

import psutil
import torch


def byte2gb(x):
    return int(x / 2**30)


class MemoryTrace:
    def __enter__(self):
        gc.collect()
        self.cuda = torch.cuda.is_available()
        if self.cuda:
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
            self.begin = byte2gb(torch.cuda.memory_allocated())
        self.process = psutil.Process()
        self.cpu_begin = byte2gb(self.process.memory_info().rss)
        return self

    def __exit__(self, *exc):
        gc.collect()
        self.peak = self.max_reserved = self.peak_active_gb = self.cuda_malloc_retires = 0
        if self.cuda:
            torch.cuda.empty_cache()
            self.peak = byte2gb(torch.cuda.max_memory_allocated())
            stats = torch.cuda.memory_stats()
            self.peak_active_gb = byte2gb(stats.get("active_bytes.all.peak", 0))
            self.cuda_malloc_retires = stats.get("num_alloc_retries", 0)
            self.max_reserved = byte2gb(torch.cuda.max_memory_reserved())
        self.cpu_end = byte2gb(self.process.memory_info().rss)
        self.cpu_peaked = max(0, self.cpu_end - self.cpu_begin)
