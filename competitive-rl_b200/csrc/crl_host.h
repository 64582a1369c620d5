// crl_host.h -- host-side helpers shared by the two ABI translation units (crl_abi.cu, crl_car_abi.cu).
#pragma once
#include <cuda_runtime.h>

// Every ABI entry point runs on the handle's device and leaves the caller's current device as it found it
// (a process that drives several GPUs must not have its current device flipped by an env on another GPU,
// including when a handle is destroyed at garbage-collection time).
struct CrlDeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit CrlDeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            switched = (err == cudaSuccess);
        }
    }
    ~CrlDeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    CrlDeviceGuard(const CrlDeviceGuard&) = delete;
    CrlDeviceGuard& operator=(const CrlDeviceGuard&) = delete;
};
