"""denormalized_b200 -- B200-native (sm_100a) windowed grouped aggregate + post-aggregate filter behind
Denormalized's operator API.  The product is the C-ABI shared library `libdnz_gpu.so` (include/dnz_gpu.h);
this package holds its sources (csrc/), the C++ host-side mirror of the reference interface (cpp/) and a thin
ctypes binding used by the tests and the benchmark.  There is no CPU fallback anywhere in this package."""
from .capi import (AGG_KINDS, DeviceBatches, ExchangeGroup, GpuStreamingWindow, DnzError, canonical_schema, lib, library_path,
                   make_record_batch)

__all__ = ["AGG_KINDS", "DeviceBatches", "ExchangeGroup", "GpuStreamingWindow", "DnzError", "canonical_schema", "lib", "library_path",
           "make_record_batch"]
