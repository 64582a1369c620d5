// Store-pattern ceiling for the Pong rasteriser: the same grid / warp-per-stack mapping / 16-byte streaming stores of
// 7056-byte frames out of shared memory, with no rasterisation work at all.  Variants: from shared memory (like the
// kernel) or straight from registers; st.global.cs or plain st.global.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o raster_store_probe raster_store_probe.cu && ./raster_store_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int DD = 84 * 84, NCH = DD / 16, FRAMES = 4;

template <bool CS> __device__ __forceinline__ void st16(uint4* p, uint4 v) {
    if (CS) asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else *p = v;
}

template <int WARPS, bool FROM_SMEM, bool CS>
__global__ void __launch_bounds__(WARPS * 32) probe(uint8_t* obs, long long n_stacks) {
    extern __shared__ uint4 smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* sm = smem + (size_t)warp * NCH;
    for (int i = lane; i < NCH; i += 32) sm[i] = make_uint4(i, i, i, i);
    __syncthreads();
    for (long long s = (long long)blockIdx.x * WARPS + warp; s < n_stacks; s += (long long)gridDim.x * WARPS) {
        uint4* out = reinterpret_cast<uint4*>(obs + (size_t)s * FRAMES * DD);
        for (int f = 0; f < FRAMES; ++f) {
#pragma unroll
            for (int j = 0; j < (NCH + 31) / 32; ++j) {
                const int c = lane + 32 * j;
                if (c < NCH) st16<CS>(out + (size_t)f * NCH + c, FROM_SMEM ? sm[c] : make_uint4(c, f, j, lane));
            }
            __syncwarp();
        }
    }
}

// plain grid-stride fill of the same buffer: the sustained pure-write rate of the device in this harness
template <bool CS>
__global__ void __launch_bounds__(256) fill(uint4* out, long long n16) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) st16<CS>(out + i, make_uint4((unsigned)i, 1u, 2u, 3u));
}

// non-persistent fills: one CTA per tile of the buffer, each thread writes PER_THREAD 16-byte vectors (strided by the
// block like torch's vectorized_elementwise_kernel); constant or index-dependent data
template <int PER_THREAD, bool VARY>
__global__ void __launch_bounds__(128) fill_tiled(uint4* out, long long n16) {
    const long long base = (long long)blockIdx.x * (128 * PER_THREAD);
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const long long i = base + k * 128 + threadIdx.x;
        if (i < n16) out[i] = VARY ? make_uint4((unsigned)i, (unsigned)(i >> 3), 2u, (unsigned)k) : make_uint4(7u, 7u, 7u, 7u);
    }
}

template <int PER_THREAD, bool VARY>
void run_fill_tiled(const char* name, uint8_t* obs, long long bytes) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const long long n16 = bytes / 16;
    const unsigned grid = (unsigned)((n16 + 128 * PER_THREAD - 1) / (128 * PER_THREAD));
    for (int i = 0; i < 5; ++i) fill_tiled<PER_THREAD, VARY><<<grid, 128>>>(reinterpret_cast<uint4*>(obs), n16);
    const int reps = 50;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) fill_tiled<PER_THREAD, VARY><<<grid, 128>>>(reinterpret_cast<uint4*>(obs), n16);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s grid %u x 128, %d vec/thread: %.3f ms/launch  %.0f GB/s  (%s)\n", name, grid, PER_THREAD, ms / reps,
           (double)bytes / (ms / reps) / 1e6, cudaGetErrorString(cudaGetLastError()));
}

template <bool CS>
void run_fill(const char* name, uint8_t* obs, long long bytes, int ctas_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    for (int i = 0; i < 5; ++i) fill<CS><<<grid, 256>>>(reinterpret_cast<uint4*>(obs), bytes / 16);
    const int reps = 50;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) fill<CS><<<grid, 256>>>(reinterpret_cast<uint4*>(obs), bytes / 16);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %d CTAs/SM x  8 warps: %.3f ms/launch  %.0f GB/s  (%s)\n", name, ctas_per_sm, ms / reps, (double)bytes / (ms / reps) / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}

// like probe<>, but the 8 warps of a CTA write ONE frame together (warp w takes chunks w, w + 8, ... of the frame's
// 32-chunk groups), so a CTA's stores in flight cover one contiguous 7 KB frame instead of eight frames 28 KB apart
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) probe_cta_frame(uint8_t* obs, long long n_stacks) {
    const long long n_frames = n_stacks * FRAMES;
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        uint4* out = reinterpret_cast<uint4*>(obs + (size_t)f * DD);
        for (int c = threadIdx.x; c < NCH; c += WARPS * 32) st16<true>(out + c, make_uint4(c, 1u, 2u, 3u));
    }
}

// persistent, warp per FRAME: the warps of a CTA write adjacent frames (a CTA round covers WARPS * 7 KB contiguous)
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) probe_warp_frame(uint8_t* obs, long long n_frames) {
    extern __shared__ uint4 smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* sm = smem + (size_t)warp * NCH;
    for (int i = lane; i < NCH; i += 32) sm[i] = make_uint4(i, i, i, i);
    __syncthreads();
    for (long long f = (long long)blockIdx.x * WARPS + warp; f < n_frames; f += (long long)gridDim.x * WARPS) {
        uint4* out = reinterpret_cast<uint4*>(obs + (size_t)f * DD);
#pragma unroll
        for (int j = 0; j < (NCH + 31) / 32; ++j) {
            const int c = lane + 32 * j;
            if (c < NCH) st16<true>(out + c, sm[c]);
        }
        __syncwarp();
    }
}

// persistent, CTA-cooperative drain: the warps fill their own buffers, then all threads of the CTA stream the WARPS
// buffers (adjacent frames) out as one contiguous region
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) probe_cta_drain(uint8_t* obs, long long n_frames) {
    extern __shared__ uint4 smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* sm = smem + (size_t)warp * NCH;
    for (int i = lane; i < NCH; i += 32) sm[i] = make_uint4(i, i, i, i);
    __syncthreads();
    for (long long f0 = (long long)blockIdx.x * WARPS; f0 < n_frames; f0 += (long long)gridDim.x * WARPS) {
        uint4* out = reinterpret_cast<uint4*>(obs + (size_t)f0 * DD);
        for (int c = threadIdx.x; c < WARPS * NCH; c += WARPS * 32) st16<true>(out + c, smem[c]);
        __syncthreads();
    }
}

// persistent, warp per stack like the kernel, but the next stack comes from a global counter (the GPU as a whole
// sweeps the buffer front to back, like the hardware CTA dispatcher does for a non-persistent launch)
template <int WARPS, bool PER_CTA, int UNIT = FRAMES>
__global__ void __launch_bounds__(WARPS * 32) probe_dynamic(uint8_t* obs, long long n_stacks_in, unsigned long long* counter) {
    const long long n_stacks = n_stacks_in * (FRAMES / UNIT);      // tickets of UNIT frames
    extern __shared__ uint4 smem[];
    __shared__ long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* sm = smem + (size_t)warp * NCH;
    for (int i = lane; i < NCH; i += 32) sm[i] = make_uint4(i, i, i, i);
    __syncthreads();
    for (;;) {
        long long s;
        if (PER_CTA) {
            if (threadIdx.x == 0) s_base = (long long)atomicAdd(counter, (unsigned long long)WARPS);
            __syncthreads();
            s = s_base + warp;
            __syncthreads();
            if (s_base >= n_stacks) break;
        } else {
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(counter, 1ull);
            s = (long long)__shfl_sync(0xffffffffu, t, 0);
        }
        if (s >= n_stacks) { if (PER_CTA) continue; else break; }
        uint4* out = reinterpret_cast<uint4*>(obs + (size_t)s * UNIT * DD);
        for (int f = 0; f < UNIT; ++f) {
#pragma unroll
            for (int j = 0; j < (NCH + 31) / 32; ++j) {
                const int c = lane + 32 * j;
                if (c < NCH) st16<true>(out + (size_t)f * NCH + c, sm[c]);
            }
            __syncwarp();
        }
    }
}

template <int WARPS, bool PER_CTA, int UNIT = FRAMES>
void run_dynamic(const char* name, uint8_t* obs, long long n_stacks, int ctas_per_sm) {
    const size_t smem = (size_t)WARPS * NCH * 16;
    cudaFuncSetAttribute(probe_dynamic<WARPS, PER_CTA, UNIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned long long* counter;
    cudaMalloc(&counter, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm, reps = 50;
    for (int i = 0; i < reps + 5; ++i) {
        if (i == 5) cudaEventRecord(e0);
        cudaMemsetAsync(counter, 0, 8);
        probe_dynamic<WARPS, PER_CTA, UNIT><<<grid, WARPS * 32, smem>>>(obs, n_stacks, counter);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %d CTAs/SM x %2d warps: %.3f ms/launch  %.0f GB/s  (%s)\n", name, ctas_per_sm, WARPS, ms / reps,
           (double)n_stacks * FRAMES * DD / (ms / reps) / 1e6, cudaGetErrorString(cudaGetLastError()));
}

// persistent, CTA ticket = WARPS / FRAMES consecutive stacks; warp w writes frame (w % FRAMES) of stack (w / FRAMES):
// the CTA's warps write one contiguous WARPS * 7 KB region at a time
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) probe_dynamic_split(uint8_t* obs, long long n_stacks, unsigned long long* counter) {
    extern __shared__ uint4 smem[];
    __shared__ long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* sm = smem + (size_t)warp * NCH;
    for (int i = lane; i < NCH; i += 32) sm[i] = make_uint4(i, i, i, i);
    __syncthreads();
    constexpr int SPC = WARPS / FRAMES;      // stacks per CTA ticket
    for (;;) {
        if (threadIdx.x == 0) s_base = (long long)atomicAdd(counter, (unsigned long long)SPC);
        __syncthreads();
        const long long s = s_base + warp / FRAMES;
        __syncthreads();
        if (s_base >= n_stacks) break;
        if (s >= n_stacks) continue;
        uint4* out = reinterpret_cast<uint4*>(obs + ((size_t)s * FRAMES + warp % FRAMES) * DD);
#pragma unroll
        for (int j = 0; j < (NCH + 31) / 32; ++j) {
            const int c = lane + 32 * j;
            if (c < NCH) st16<true>(out + c, sm[c]);
        }
    }
}

template <int WARPS>
void run_dynamic_split(const char* name, uint8_t* obs, long long n_stacks, int ctas_per_sm) {
    const size_t smem = (size_t)WARPS * NCH * 16;
    cudaFuncSetAttribute(probe_dynamic_split<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned long long* counter;
    cudaMalloc(&counter, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm, reps = 50;
    for (int i = 0; i < reps + 5; ++i) {
        if (i == 5) cudaEventRecord(e0);
        cudaMemsetAsync(counter, 0, 8);
        probe_dynamic_split<WARPS><<<grid, WARPS * 32, smem>>>(obs, n_stacks, counter);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %d CTAs/SM x %2d warps: %.3f ms/launch  %.0f GB/s  (%s)\n", name, ctas_per_sm, WARPS, ms / reps,
           (double)n_stacks * FRAMES * DD / (ms / reps) / 1e6, cudaGetErrorString(cudaGetLastError()));
}

template <int WARPS, int KIND>
void run_frames(const char* name, uint8_t* obs, long long n_stacks, int ctas_per_sm) {
    const size_t smem = (size_t)WARPS * NCH * 16;
    if (KIND == 0) cudaFuncSetAttribute(probe_warp_frame<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else cudaFuncSetAttribute(probe_cta_drain<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    const int reps = 50;
    for (int i = 0; i < reps + 5; ++i) {
        if (i == 5) cudaEventRecord(e0);
        if (KIND == 0) probe_warp_frame<WARPS><<<grid, WARPS * 32, smem>>>(obs, n_stacks * FRAMES);
        else probe_cta_drain<WARPS><<<grid, WARPS * 32, smem>>>(obs, n_stacks * FRAMES);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %d CTAs/SM x %2d warps: %.3f ms/launch  %.0f GB/s  (%s)\n", name, ctas_per_sm, WARPS, ms / reps,
           (double)n_stacks * FRAMES * DD / (ms / reps) / 1e6, cudaGetErrorString(cudaGetLastError()));
}

template <int WARPS>
void run_cta_frame(const char* name, uint8_t* obs, long long n_stacks, int ctas_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    for (int i = 0; i < 5; ++i) probe_cta_frame<WARPS><<<grid, WARPS * 32>>>(obs, n_stacks);
    const int reps = 50;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) probe_cta_frame<WARPS><<<grid, WARPS * 32>>>(obs, n_stacks);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %d CTAs/SM x %2d warps: %.3f ms/launch  %.0f GB/s  (%s)\n", name, ctas_per_sm, WARPS, ms / reps,
           (double)n_stacks * FRAMES * DD / (ms / reps) / 1e6, cudaGetErrorString(cudaGetLastError()));
}

template <int WARPS, bool FROM_SMEM, bool CS>
void run(const char* name, uint8_t* obs, long long n_stacks, int ctas_per_sm) {
    const size_t smem = (size_t)WARPS * NCH * 16;
    cudaFuncSetAttribute(probe<WARPS, FROM_SMEM, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    for (int i = 0; i < 5; ++i) probe<WARPS, FROM_SMEM, CS><<<grid, WARPS * 32, smem>>>(obs, n_stacks);
    const int reps = 50;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) probe<WARPS, FROM_SMEM, CS><<<grid, WARPS * 32, smem>>>(obs, n_stacks);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)n_stacks * FRAMES * DD;
    printf("%-44s %d CTAs/SM x %2d warps: %.3f ms/launch  %.0f GB/s  (%s)\n", name, ctas_per_sm, WARPS, ms / reps,
           bytes / (ms / reps) / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const long long n_stacks = 65536LL * 2;
    uint8_t* obs;
    cudaMalloc(&obs, (size_t)n_stacks * FRAMES * DD);
    run<8, true, true>("smem -> st.global.cs (the kernel's drain)", obs, n_stacks, 3);
    run<8, true, false>("smem -> st.global", obs, n_stacks, 3);
    run<8, false, true>("registers -> st.global.cs", obs, n_stacks, 3);
    run<8, false, false>("registers -> st.global", obs, n_stacks, 3);
    run<8, false, false>("registers -> st.global", obs, n_stacks, 8);
    run<14, true, true>("smem -> st.global.cs", obs, n_stacks, 2);
    run<4, true, true>("smem -> st.global.cs", obs, n_stacks, 7);
    run_fill<false>("grid-stride fill, st.global", obs, n_stacks * FRAMES * DD, 8);
    run_fill<true>("grid-stride fill, st.global.cs", obs, n_stacks * FRAMES * DD, 8);
    run_fill<false>("grid-stride fill, st.global", obs, n_stacks * FRAMES * DD, 4);
    run_fill_tiled<1, false>("tiled fill, constant data", obs, n_stacks * FRAMES * DD);
    run_fill_tiled<4, false>("tiled fill, constant data", obs, n_stacks * FRAMES * DD);
    run_fill_tiled<1, true>("tiled fill, varying data", obs, n_stacks * FRAMES * DD);
    run_fill_tiled<4, true>("tiled fill, varying data", obs, n_stacks * FRAMES * DD);
    run_fill_tiled<8, true>("tiled fill, varying data", obs, n_stacks * FRAMES * DD);
    run_dynamic<8, false>("persistent, next stack from a counter (warp)", obs, n_stacks, 3);
    run_dynamic<8, true>("persistent, next 8 stacks from a counter (CTA)", obs, n_stacks, 3);
    run_dynamic_split<8>("CTA ticket = 2 stacks, warp = one frame", obs, n_stacks, 3);
    run_dynamic_split<8>("CTA ticket = 2 stacks, warp = one frame", obs, n_stacks, 2);
    run_dynamic_split<8>("CTA ticket = 2 stacks, warp = one frame", obs, n_stacks, 4);
    run_dynamic_split<4>("CTA ticket = 1 stack, warp = one frame", obs, n_stacks, 7);
    run_dynamic_split<16>("CTA ticket = 4 stacks, warp = one frame", obs, n_stacks, 1);
    run_dynamic<8, false, 1>("persistent, next FRAME from a counter (warp)", obs, n_stacks, 3);
    run_dynamic<8, false, 2>("persistent, next 2 frames from a counter", obs, n_stacks, 3);
    run_dynamic<8, false>("persistent, next stack from a counter (warp)", obs, n_stacks, 4);
    run_dynamic<8, false>("persistent, next stack from a counter (warp)", obs, n_stacks, 2);
    run_dynamic<4, false>("persistent, next stack from a counter (warp)", obs, n_stacks, 7);
    run_dynamic<4, false>("persistent, next stack from a counter (warp)", obs, n_stacks, 5);
    run_frames<8, 0>("persistent, warp per frame, adjacent frames", obs, n_stacks, 3);
    run_frames<8, 1>("persistent, CTA-cooperative drain of 8 frames", obs, n_stacks, 3);
    run_frames<4, 1>("persistent, CTA-cooperative drain of 4 frames", obs, n_stacks, 7);
    run_cta_frame<8>("CTA writes one frame at a time (cs)", obs, n_stacks, 8);
    run_cta_frame<4>("CTA writes one frame at a time (cs)", obs, n_stacks, 16);
    return 0;
}
