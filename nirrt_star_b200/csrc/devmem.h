// devmem.h -- device allocations of libnirrt_b200.so.
// A planner batch or a network engine is ~60 cudaMalloc calls to create and as many cudaFree calls to destroy, and the
// stand-alone predicates allocate temporaries per call; on a host that shares the driver with other tenants those calls take
// anywhere from 0.1 s to 1.5 s per create / destroy pair.  Freed blocks are therefore kept in a process-wide cache (exact size
// match per device, capped) and handed out again zero-filled.  nirrt_dev_free keeps cudaFree's contract of returning only
// when the device is idle, which callers rely on for temporaries used by asynchronous work.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

cudaError_t nirrt_dev_malloc(void **p, size_t bytes);      // on the current device
void nirrt_dev_free(void *p);
