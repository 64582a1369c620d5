// pong_raster.cu -- frame rasteriser fused with the atari_wrappers preprocessing.
//
// Replaces, in ONE kernel and without ever materialising a 210x160x3 frame
// (paths relative to /root/reference/competitive_rl/):
//   _get_screen_img / _get_screen_img_double_player   pong/base_pong_env.py:66-74, 149-155
//   Arena.draw, Ball.draw, Bat.draw, Scoreboard.draw   pong/base_pong_env.py:278-280, 322-323, 406-410, 480-487
//   MaxAndSkipEnv max over the two buffered frames     utils/atari_wrappers.py:153-156
//   WarpFrame: cv2.cvtColor(RGB2GRAY) + cv2.resize(INTER_AREA)   utils/atari_wrappers.py:215-219
//   FrameStack._get_ob + WrapPyTorch (CHW, oldest first)         utils/atari_wrappers.py:257-259, 35-37
//
// Arithmetic restated from cv2 4.13 (pinned in tests/test_oracle_cv2.py):
//   gray  = (R*9798 + G*19235 + B*3735 + 16384) >> 15
//   area  = per source row: buf = sum_t fl(fl(S)*alpha_t) (ascending sx, un-fused);
//           rows: sum = fl(beta_0*buf_0), then sum = fl(sum + fl(beta_k*buf_k));
//           dst = saturate_u8(rint_half_even(sum)).   __fmul_rn/__fadd_rn keep it un-fused.
//
// Structure exploited: a Pong frame is white border + black arena + <=3 white
// rectangles + the scoreboard text above the arena.  So the preprocessed frame is
//   * rows above the arena: a function of the score pair(s) only -> looked up from a
//     table built once per atlas (pong_build_tables_kernel),
//   * everything else: a constant template, except the few destination pixels whose
//     3x3 (5x5 at 42x42) source footprint touches a rectangle -> evaluated exactly.
// Each warp stages one frame in shared memory (template fill, pixel patches), then
// streams it out with coalesced 16-byte stores.
#include "pong_common.cuh"

namespace crl {

// ---------------------------------------------------------------------------------
// Per-frame context: what the two max-pooled rendered frames contain, in the
// viewing agent's coordinates (agent 1 sees rows >= 25 mirrored in x).
struct FrameCtx {
    uint32_t rect[6];   // x0 | x1<<8 | y0<<16 | y1<<24 (arena-clipped; 0 = empty)
    int pairA, pairB;   // atlas score-pair index l*22+r of the two frames
    bool mirror;
    bool any_valid;
};

__device__ __forceinline__ uint32_t pack_rect(int x0, int x1, int y0, int y1, bool mirror) {
    x0 = max(x0, 0); x1 = min(x1, SCREEN_W);
    y0 = max(y0, ARENA_TOP); y1 = min(y1, ARENA_BOTTOM);   // white-on-white outside the arena
    if (x1 <= x0 || y1 <= y0) return 0u;
    if (mirror) { const int t = SCREEN_W - x1; x1 = SCREEN_W - x0; x0 = t; }
    return (uint32_t)x0 | ((uint32_t)x1 << 8) | ((uint32_t)y0 << 16) | ((uint32_t)y1 << 24);
}

__device__ __forceinline__ FrameCtx make_ctx(const FrameSpec f, const int agent) {
    FrameCtx c;
    uint32_t ax = f.x, ay = f.y, bx = f.z, by = f.w;
    const bool va = (ay >> 16) & 1u, vb = (by >> 16) & 1u;
    c.any_valid = va || vb;
    c.mirror = agent != 0;
    if (!va) { ax = bx; ay = by; }      // max(0-frame, X) = X
    if (!vb) { bx = ax; by = ay; }
    c.pairA = (int)(ay & 255u) * ATLAS_SCORES + (int)((ay >> 8) & 255u);
    c.pairB = (int)(by & 255u) * ATLAS_SCORES + (int)((by >> 8) & 255u);
    {
        const int x = ax & 255u, y = (ax >> 8) & 255u, l = (ax >> 16) & 255u, r = (ax >> 24) & 255u;
        c.rect[0] = pack_rect(x, x + BALL_SIZE, y, y + BALL_SIZE, c.mirror);
        c.rect[1] = pack_rect(LEFT_BAT_X, LEFT_BAT_X + BAT_W, l, l + BAT_H, c.mirror);
        c.rect[2] = pack_rect(RIGHT_BAT_X, RIGHT_BAT_X + BAT_W, r, r + BAT_H, c.mirror);
    }
    if (ax == bx) {
        c.rect[3] = c.rect[4] = c.rect[5] = 0u;
    } else {
        const int x = bx & 255u, y = (bx >> 8) & 255u, l = (bx >> 16) & 255u, r = (bx >> 24) & 255u;
        c.rect[3] = pack_rect(x, x + BALL_SIZE, y, y + BALL_SIZE, c.mirror);
        c.rect[4] = pack_rect(LEFT_BAT_X, LEFT_BAT_X + BAT_W, l, l + BAT_H, c.mirror);
        c.rect[5] = pack_rect(RIGHT_BAT_X, RIGHT_BAT_X + BAT_W, r, r + BAT_H, c.mirror);
    }
    return c;
}

// bit t set <=> lo <= s0 + t < hi, for t in [0, n)
__device__ __forceinline__ uint32_t span_bits(int lo, int hi, int s0, int n) {
    const int a = max(lo - s0, 0), b = min(hi - s0, n);
    return (b > a) ? (((1u << b) - 1u) & ~((1u << a) - 1u)) : 0u;
}

// gray value of the max-pooled source pixel in the rows above the arena
__device__ __forceinline__ int text_gray(const FrameCtx& c, const uint8_t* __restrict__ atlas, int sy, int sx) {
    const int ax = (c.mirror && sy >= MIRROR_ROW) ? (SCREEN_W - 1 - sx) : sx;
    const uint8_t* pa = atlas + ((size_t)(c.pairA * ATLAS_ROWS + sy) * SCREEN_W + ax) * 3;
    const uint8_t* pb = atlas + ((size_t)(c.pairB * ATLAS_ROWS + sy) * SCREEN_W + ax) * 3;
    const int r = max((int)pa[0], (int)pb[0]), g = max((int)pa[1], (int)pb[1]), b = max((int)pa[2], (int)pb[2]);
    return (r * 9798 + g * 19235 + b * 3735 + 16384) >> 15;   // cv2 RGB2GRAY, 15-bit fixed point
}

// Exact value of destination pixel (dy, dx): cv2's INTER_AREA float path evaluated on
// the analytically-described source frame.
__device__ __forceinline__ uint8_t eval_pixel(const AreaTabs* __restrict__ T, const FrameCtx& c,
                                              const uint8_t* __restrict__ atlas, int dy, int dx) {
    const int sx0 = T->x_src0[dx], nx = T->x_n[dx];
    const int sy0 = T->y_src0[dy], ny = T->y_n[dy];
    uint32_t hp[6], vp[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const uint32_t r = c.rect[k];
        hp[k] = span_bits(r & 255u, (r >> 8) & 255u, sx0, nx);
        vp[k] = span_bits((r >> 16) & 255u, r >> 24, sy0, ny);
    }
    float sum = 0.f;
#pragma unroll 1
    for (int ty = 0; ty < ny; ++ty) {
        const int sy = sy0 + ty;
        float buf = 0.f;
        if (sy < ARENA_TOP) {
            for (int tx = 0; tx < nx; ++tx)
                buf = __fadd_rn(buf, __fmul_rn((float)text_gray(c, atlas, sy, sx0 + tx), T->x_a[dx][tx]));
        } else {
            uint32_t pat = 0u;
            if (sy >= ARENA_BOTTOM) {
                pat = 0xffffffffu;
            } else {
#pragma unroll
                for (int k = 0; k < 6; ++k) pat |= ((vp[k] >> ty) & 1u) ? hp[k] : 0u;
            }
#pragma unroll
            for (int tx = 0; tx < MAX_TAPS; ++tx)
                if (tx < nx && ((pat >> tx) & 1u)) buf = __fadd_rn(buf, T->x_pa[dx][tx]);
        }
        const float term = __fmul_rn(T->y_b[dy][ty], buf);
        sum = (ty == 0) ? term : __fadd_rn(sum, term);
    }
    const int v = __float2int_rn(sum);   // cvRound: round half to even
    return (uint8_t)min(max(v, 0), 255);
}

// ---------------------------------------------------------------------------------
// Table builder: text_tab[(pair*3+kind)*2+agent][text_stride] and tmpl[dim*dim (+pad)].
// kind 0: both frames show `pair`; 1: the other frame shows (l+1, r); 2: (l, r+1).
__global__ void pong_build_tables_kernel(PongDev p, uint8_t* text_tab, uint8_t* tmpl, int tmpl_bytes) {
    const int dd = p.dim * p.dim;
    const int entries = ATLAS_SCORES * ATLAS_SCORES * 3 * 2;
    const long long total = (long long)entries * p.text_stride + tmpl_bytes;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        FrameCtx c;
#pragma unroll
        for (int k = 0; k < 6; ++k) c.rect[k] = 0u;
        c.any_valid = true;
        int pix;
        uint8_t* dst;
        if (i < (long long)entries * p.text_stride) {
            const int entry = (int)(i / p.text_stride);
            pix = (int)(i % p.text_stride);
            const int agent = entry & 1, kind = (entry >> 1) % 3, pair = (entry >> 1) / 3;
            int l = pair / ATLAS_SCORES, r = pair % ATLAS_SCORES;
            c.pairA = pair;
            if (kind == 1 && l + 1 < ATLAS_SCORES) l += 1;
            if (kind == 2 && r + 1 < ATLAS_SCORES) r += 1;
            c.pairB = l * ATLAS_SCORES + r;
            c.mirror = agent != 0;
            dst = text_tab + i;
        } else {
            pix = (int)(i - (long long)entries * p.text_stride);
            c.pairA = c.pairB = 0;
            c.mirror = false;
            dst = tmpl + pix;
        }
        *dst = (pix < dd) ? eval_pixel(p.tabs, c, p.atlas, pix / p.dim, pix % p.dim) : (uint8_t)0;
    }
}

// ---------------------------------------------------------------------------------
// Reference rasteriser: one thread per destination pixel, no tables, no shortcuts.
// Used for terminal observations (rare) and as the in-library cross-check of the
// fast kernel (tests/test_gpu_pong_parity.py).
__global__ void pong_raster_generic_kernel(PongDev p, const FrameSpec* __restrict__ hist,
                                           const uint8_t* __restrict__ only_done, uint8_t* obs0, uint8_t* obs1) {
    const int dd = p.dim * p.dim;
    const long long total = (long long)p.n * p.n_agents * p.c * dd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int pix = (int)(i % dd);
        long long f = i / dd;
        const int slot = (int)(f % p.c);
        f /= p.c;
        const int env = (int)(f % p.n), agent = (int)(f / p.n);
        if (only_done != nullptr && !only_done[env]) continue;
        const FrameCtx c = make_ctx(hist[(size_t)slot * p.n + env], agent);
        uint8_t* out = (agent ? obs1 : obs0) + ((size_t)env * p.c + slot) * dd;
        out[pix] = c.any_valid ? eval_pixel(p.tabs, c, p.atlas, pix / p.dim, pix % p.dim) : (uint8_t)0;
    }
}

// ---------------------------------------------------------------------------------
// Hot kernel: one warp per (agent, env, stack slot) frame.
template <int VEC> struct VecT;
template <> struct VecT<16> { typedef uint4 type; };
template <> struct VecT<4> { typedef uint32_t type; };

__device__ __forceinline__ void st_stream(uint4* ptr, const uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream(uint32_t* ptr, const uint32_t v) {
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}

constexpr int RASTER_WARPS = 8;

template <int VEC>
__global__ void __launch_bounds__(RASTER_WARPS * 32)
pong_raster_kernel(PongDev p, const FrameSpec* __restrict__ hist, uint8_t* __restrict__ obs0,
                   uint8_t* __restrict__ obs1, int frame_smem_bytes) {
    typedef typename VecT<VEC>::type V;
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long frame = (long long)blockIdx.x * RASTER_WARPS + warp;
    const long long n_frames = (long long)p.n * p.n_agents * p.c;
    if (frame >= n_frames) return;   // whole warp exits together
    const int slot = (int)(frame % p.c);
    const long long ae = frame / p.c;
    const int env = (int)(ae % p.n), agent = (int)(ae / p.n);
    const int dim = p.dim, dd = dim * dim;
    const int n_chunks = dd / VEC;
    uint8_t* sm8 = reinterpret_cast<uint8_t*>(smem_raw) + (size_t)warp * frame_smem_bytes;
    V* sm = reinterpret_cast<V*>(sm8);
    V* out = reinterpret_cast<V*>((agent ? obs1 : obs0) + ((size_t)env * p.c + slot) * dd);

    const FrameSpec spec = hist[(size_t)slot * p.n + env];
    const FrameCtx c = make_ctx(spec, agent);
    const AreaTabs* __restrict__ T = p.tabs;

    if (!c.any_valid) {   // both MaxAndSkip buffers still np.zeros
        V z;
        memset(&z, 0, sizeof z);
        for (int k = lane; k < n_chunks; k += 32) st_stream(out + k, z);
        return;
    }

    // ---- which precomputed scoreboard rows apply ----
    int base = c.pairA, kind = 0;
    bool text_ok = true;
    if (c.pairA != c.pairB) {
        const int d = c.pairB - c.pairA;   // +22: left scored, +1: right scored (one point per env-step at most)
        if (d == ATLAS_SCORES) kind = 1;
        else if (d == 1 && (c.pairA % ATLAS_SCORES) != ATLAS_SCORES - 1) kind = 2;
        else if (d == -ATLAS_SCORES) { base = c.pairB; kind = 1; }
        else if (d == -1 && (c.pairB % ATLAS_SCORES) != ATLAS_SCORES - 1) { base = c.pairB; kind = 2; }
        else text_ok = false;
    }
    const int text_chunks = p.text_stride / VEC;
    const V* __restrict__ te =
        reinterpret_cast<const V*>(p.text_tab + (size_t)((base * 3 + kind) * 2 + agent) * p.text_stride);
    const V* __restrict__ tm = reinterpret_cast<const V*>(p.tmpl);

    // ---- 1. fill the staged frame: scoreboard rows from the table, the rest from the template ----
    for (int k = lane; k < n_chunks; k += 32) sm[k] = (k < text_chunks) ? te[k] : tm[k];

    // ---- 2. patch the destination pixels whose footprint touches a rectangle ----
    // groups: left bats (A u B), right bats (A u B), ball A, ball B, [all text rows if the
    // score combination is not in the table]
    int gx0[5], gw[5], gy0[5], cum[5];
    int total = 0;
#pragma unroll
    for (int g = 0; g < 5; ++g) {
        uint32_t r0, r1 = 0u;
        if (g == 0) { r0 = c.rect[1]; r1 = c.rect[4]; }
        else if (g == 1) { r0 = c.rect[2]; r1 = c.rect[5]; }
        else if (g == 2) r0 = c.rect[0];
        else if (g == 3) r0 = c.rect[3];
        else r0 = 0u;
        int w = 0, h = 0, x0 = 0, y0 = 0;
        if (g == 4) {
            if (!text_ok) { w = dim; h = T->text_rows; }
        } else {
            if (r0 == 0u) { r0 = r1; r1 = 0u; }
            if (r0 != 0u) {
                int sx0 = r0 & 255u, sx1 = (r0 >> 8) & 255u, sy0 = (r0 >> 16) & 255u, sy1 = r0 >> 24;
                if (r1 != 0u) {
                    sx0 = min(sx0, (int)(r1 & 255u)); sx1 = max(sx1, (int)((r1 >> 8) & 255u));
                    sy0 = min(sy0, (int)((r1 >> 16) & 255u)); sy1 = max(sy1, (int)(r1 >> 24));
                }
                x0 = T->x_first[sx0];
                w = T->x_last[sx1 - 1] + 1 - x0;
                y0 = T->y_first[sy0];
                h = T->y_last[sy1 - 1] + 1 - y0;
            }
        }
        gx0[g] = x0; gw[g] = w; gy0[g] = y0;
        total += w * h;
        cum[g] = total;
    }
    __syncwarp();
    for (int q = lane; q < total; q += 32) {
        int g = 0, start = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (q >= cum[k]) { g = k + 1; start = cum[k]; }
        int x0 = gx0[0], w = gw[0], y0 = gy0[0];
#pragma unroll
        for (int k = 1; k < 5; ++k)
            if (g == k) { x0 = gx0[k]; w = gw[k]; y0 = gy0[k]; }
        const int local = q - start;
        const int dy = y0 + local / w, dx = x0 + local % w;
        sm8[dy * dim + dx] = eval_pixel(T, c, p.atlas, dy, dx);
    }
    __syncwarp();

    // ---- 3. stream the frame out: coalesced VEC-byte stores ----
    for (int k = lane; k < n_chunks; k += 32) st_stream(out + k, sm[k]);
}

// ---------------------------------------------------------------------------------
// Raw 210x160x3 export of one env's CURRENT game state (VecEnv.render / get_images,
// utils/base_vec_env.py:189-216).  Not on the step path.
__global__ void pong_raw_frame_kernel(PongDev p, int env, uint8_t* rgb0, uint8_t* rgb1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= SCREEN_W * SCREEN_H) return;
    const int y = i / SCREEN_W, x = i % SCREEN_W;
    const int32_t b = p.ball[env], t = p.bats[env], s = p.score[env];
    const int bx = (int16_t)(b & 0xffff), by = (int16_t)(b >> 16), ly = t & 0xffff, ry = (t >> 16) & 0xffff;
    const int pair = (s & 0xff) * ATLAS_SCORES + ((s >> 8) & 0xff);
    uint8_t r, g, bl;
    if (y < ATLAS_ROWS) {
        const uint8_t* a = p.atlas + ((size_t)(pair * ATLAS_ROWS + y) * SCREEN_W + x) * 3;
        r = a[0]; g = a[1]; bl = a[2];
    } else if (y >= ARENA_BOTTOM) {
        r = g = bl = 255;
    } else {
        const bool w = (x >= bx && x < bx + BALL_SIZE && y >= by && y < by + BALL_SIZE) ||
                       (x >= LEFT_BAT_X && x < LEFT_BAT_X + BAT_W && y >= ly && y < ly + BAT_H) ||
                       (x >= RIGHT_BAT_X && x < RIGHT_BAT_X + BAT_W && y >= ry && y < ry + BAT_H);
        r = g = bl = w ? 255 : 0;
    }
    uint8_t* o0 = rgb0 + (size_t)i * 3;
    o0[0] = r; o0[1] = g; o0[2] = bl;
    if (rgb1 != nullptr) {
        const int xm = (y >= MIRROR_ROW) ? SCREEN_W - 1 - x : x;
        uint8_t* o1 = rgb1 + ((size_t)y * SCREEN_W + xm) * 3;
        o1[0] = r; o1[1] = g; o1[2] = bl;
    }
}

// ---------------------------------------------------------------------------------
static inline int frame_smem_bytes(const PongDev& p) { return ((p.dim * p.dim + 15) / 16) * 16; }

// per-device opt-in to >48 KB dynamic shared memory (8 warps x 7056 B at 84x84)
cudaError_t pong_raster_init() {
    const int bytes = RASTER_WARPS * (((MAX_DIM * MAX_DIM + 15) / 16) * 16);
    cudaError_t e = cudaFuncSetAttribute(pong_raster_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(pong_raster_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

cudaError_t launch_pong_build_tables(const PongDev& p, uint8_t* text_tab, uint8_t* tmpl, cudaStream_t s) {
    pong_build_tables_kernel<<<148 * 8, 256, 0, s>>>(p, text_tab, tmpl, frame_smem_bytes(p));
    return cudaGetLastError();
}

cudaError_t launch_pong_raster_generic(const PongDev& p, const FrameSpec* hist, const uint8_t* only_done,
                                       uint8_t* obs0, uint8_t* obs1, cudaStream_t s) {
    pong_raster_generic_kernel<<<148 * 16, 256, 0, s>>>(p, hist, only_done, obs0, obs1);
    return cudaGetLastError();
}

cudaError_t launch_pong_raster(const PongDev& p, const FrameSpec* hist, uint8_t* obs0, uint8_t* obs1, cudaStream_t s) {
    const long long n_frames = (long long)p.n * p.n_agents * p.c;
    if (n_frames == 0) return cudaSuccess;
    const int fsb = frame_smem_bytes(p);
    const size_t smem = (size_t)RASTER_WARPS * fsb;
    const unsigned grid = (unsigned)((n_frames + RASTER_WARPS - 1) / RASTER_WARPS);
    const bool vec16 = (p.dim * p.dim) % 16 == 0 && ((uintptr_t)obs0 % 16 == 0) && ((uintptr_t)obs1 % 16 == 0);
    if (vec16)
        pong_raster_kernel<16><<<grid, RASTER_WARPS * 32, smem, s>>>(p, hist, obs0, obs1, fsb);
    else
        pong_raster_kernel<4><<<grid, RASTER_WARPS * 32, smem, s>>>(p, hist, obs0, obs1, fsb);
    return cudaGetLastError();
}

cudaError_t launch_pong_raw_frame(const PongDev& p, int env, uint8_t* rgb0, uint8_t* rgb1, cudaStream_t s) {
    pong_raw_frame_kernel<<<(SCREEN_W * SCREEN_H + 255) / 256, 256, 0, s>>>(p, env, rgb0, rgb1);
    return cudaGetLastError();
}

}  // namespace crl
