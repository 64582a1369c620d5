// car_raster.cu -- 96x96 grayscale observation of cCarRacing, one CTA per (env, player) frame.
//
// Replaces (paths relative to /root/reference/competitive_rl/): get_observation (car_racing_multi_players.py
// :622-634), camera_update (:791-812), camera_view (:764-789), render_road_for_observation_map (:732-755),
// render(mode="internal_rgb_array") (:857-863), Car.draw_for_pygame (car_dynamics.py:284-298),
// render_indicators_for_pygame (:645-670) + pygame_rendering.py:8-18, and the FrameStack /
// MultipleFrameStack + FlattenMultiAgentObservation + WrapPyTorch layout (utils/atari_wrappers.py:222-334).
//
// The reference pre-renders a 10 000 x 10 000 px road map per reset (400 MB surface), crops 192x192,
// rotates and centre-blits.  Here the frame is rasterised directly: the CTA culls the road tiles near
// the camera, projects their polygons to screen space and fills them into a shared-memory key buffer
// with atomicMax on (paint order << 8 | gray), so all polygons are filled in parallel yet the
// reference's painter's order decides every pixel.  Sampling rule: colour of the analytic scene at the
// destination pixel centre (same rule as oracle/car_oracle.c; pygame's scan conversion and rotozoom
// cannot be reproduced without pygame).  Gray values come from a host-computed palette
// (trunc(0.299 R + 0.587 G + 0.114 B) in fp64 like :632-633).
#include <math.h>

#include "car_common.cuh"

namespace crl {

#define CR_SIZE 0.02
#define CR_PLAYFIELD (2000.0 / 6.0)
#define CR_TRACK_WIDTH (40.0 / 6.0)
#define CR_BORDER (8.0 / 6.0)
#define CR_TRACK_DETAIL_STEP (21.0 / 6.0)

constexpr int RASTER_THREADS = 256;
constexpr int MAX_CAND = 160;

__constant__ float c_hull_poly[4][8][2] = {
    {{-60, +130}, {+60, +130}, {+60, +110}, {-60, +110}},
    {{-15, +120}, {+15, +120}, {+20, +20}, {-20, 20}},
    {{+25, +20}, {+50, -10}, {+50, -40}, {+20, -90}, {-20, -90}, {-50, -40}, {-50, -10}, {-25, +20}},
    {{-50, -120}, {+50, -120}, {+50, -90}, {-50, -90}}};
__constant__ int c_hull_count[4] = {4, 4, 8, 4};
__constant__ float c_wheelpos_r[4][2] = {{-55, +80}, {+55, +80}, {-55, -82}, {+55, -82}};

struct Cam {
    double camx, camy, s_rot, c_rot, k;   // k = obs_scale px per world unit
};

// world -> screen (pixel coordinates, y down, origin at the top-left corner)
__device__ __forceinline__ void to_screen(const Cam& cm, double wx, double wy, float& sx, float& sy) {
    const double mx = (cm.camx - wx) * cm.k, my = (cm.camy - wy) * cm.k;     // map-pixel offset from the camera
    sx = (float)(mx * cm.c_rot + my * cm.s_rot + CAR_W / 2.0);               // rotate CCW by the camera angle
    sy = (float)(-mx * cm.s_rot + my * cm.c_rot + CAR_H / 2.0);
}

constexpr int KEY_STRIDE = CAR_W + 1;   // padded row stride: lanes work on different rows of the same columns

// Fill a convex polygon given in screen space into the key buffer: the warp sweeps the polygon's
// bounding box (32 pixels per pass, shaped 1x32, 2x16 or 4x8 to suit the box width) and every lane
// evaluates the edge functions at its pixel centre.  Covered = all cross products >= 0 (or all <= 0).
template <int N>
__device__ __forceinline__ void fill_poly_n(unsigned int* keys, const float* sx, const float* sy, unsigned int key, int lane) {
    float minx = sx[0], maxx = sx[0], miny = sy[0], maxy = sy[0], area2 = 0.f;
    float ex[N], ey[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int i2 = (i + 1 == N) ? 0 : i + 1;
        minx = fminf(minx, sx[i]); maxx = fmaxf(maxx, sx[i]); miny = fminf(miny, sy[i]); maxy = fmaxf(maxy, sy[i]);
        area2 += sx[i] * sy[i2] - sx[i2] * sy[i];
        ex[i] = sx[i2] - sx[i]; ey[i] = sy[i2] - sy[i];
    }
    const int y0 = max(0, (int)floorf(miny - 0.5f)), y1 = min(CAR_H - 1, (int)ceilf(maxy - 0.5f));
    const int x0 = max(0, (int)floorf(minx - 0.5f)), x1 = min(CAR_W - 1, (int)ceilf(maxx - 0.5f));
    if (x1 < x0 || y1 < y0) return;
    const bool ccw = area2 >= 0.f;
    const int bw = x1 - x0 + 1;
    const int lw = (bw <= 8) ? 3 : (bw <= 16) ? 4 : 5;          // log2 of the pass width
    const int lx = lane & ((1 << lw) - 1), ly = lane >> lw, rows_per_pass = 32 >> lw;
    for (int yb = y0; yb <= y1; yb += rows_per_pass) {
        const int y = yb + ly;
        const float cy = y + 0.5f;
        for (int xb = x0; xb <= x1; xb += (1 << lw)) {
            const int x = xb + lx;
            const float cx = x + 0.5f;
            bool pos = false, neg = false;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const float cr = ex[i] * (cy - sy[i]) - ey[i] * (cx - sx[i]);
                pos = pos || cr > 0.f;
                neg = neg || cr < 0.f;
            }
            if ((ccw ? !neg : !pos) && y <= y1 && x <= x1) atomicMax(&keys[y * KEY_STRIDE + x], key);
        }
    }
}

__device__ __forceinline__ void fill_poly(unsigned int* keys, const float* sx, const float* sy, int n, unsigned int key, int lane) {
    switch (n) {
        case 3: fill_poly_n<3>(keys, sx, sy, key, lane); break;
        case 4: fill_poly_n<4>(keys, sx, sy, key, lane); break;
        case 5: fill_poly_n<5>(keys, sx, sy, key, lane); break;
        case 8: fill_poly_n<8>(keys, sx, sy, key, lane); break;
        default: break;
    }
}

__device__ __forceinline__ void hud_rect(uint8_t* img, double x, double y, double w, double h, uint8_t val, int tid) {
    // pygame.draw.rect((x, y, w, h)): Rect truncates each float; a negative extent grows the other way
    const int X = (int)x, Y = (int)y, W = (int)w, H = (int)h;
    if (W == 0 || H == 0) return;
    int x0 = W >= 0 ? X : X + W, x1 = W >= 0 ? X + W : X + 1;
    int y0 = H >= 0 ? Y : Y + H, y1 = H >= 0 ? Y + H : Y + 1;
    x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, CAR_W); y1 = min(y1, CAR_H);
    const int bw = x1 - x0, total = bw * (y1 - y0);
    if (bw <= 0 || total <= 0) return;
    for (int q = tid; q < total; q += RASTER_THREADS) img[(y0 + q / bw) * CAR_W + x0 + q % bw] = val;
}

__global__ void __launch_bounds__(RASTER_THREADS)
car_render_kernel(CarDev p, int only_done, uint8_t* __restrict__ obs, uint8_t* __restrict__ term_obs) {
    __shared__ unsigned int keys[CAR_H * KEY_STRIDE];
    __shared__ __align__(16) uint8_t img[CAR_PIX];
    __shared__ int cand[MAX_CAND];
    __shared__ int n_cand;
    __shared__ Cam cam;
    __shared__ float car_body[CAR_MAX_PLAYERS][40];
    __shared__ double hud_vals[8];

    const int frame = blockIdx.x;                     // env * players + player
    const int e = frame / p.players, pi = frame % p.players;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (only_done && !p.env_done[e]) return;
    const CarHullConst* K = p.consts;
    const uint8_t* G = K->gray;

    // ---- camera (camera_update "rgb_array": hull.position + R(angle) * (0, 16)) and car states ----
    if (tid < p.players * 40) car_body[tid / 40][tid % 40] = p.body[((size_t)e * p.players + tid / 40) * 40 + tid % 40];
    __syncthreads();
    if (tid == 0) {
        const float* b = car_body[pi];
        float hs, hc;
        sincosf(b[2], &hs, &hc);
        const float hx = b[0] - (hc * K->hull_lcx - hs * K->hull_lcy), hy = b[1] - (hs * K->hull_lcx + hc * K->hull_lcy);
        double angle = (double)b[2];
        const double vx = (double)b[3], vy = (double)b[4];
        if (vx * vx + vy * vy > 0.5 * 0.5) angle = atan2(-vx, +vy);
        const float fa = (float)angle;
        float fs, fc;
        sincosf(fa, &fs, &fc);
        cam.camx = (double)hx + (double)(fc * 0.0f - fs * 16.0f);
        cam.camy = (double)hy + (double)(fs * 0.0f + fc * 16.0f);
        cam.s_rot = sin(angle);
        cam.c_rot = cos(angle);
        cam.k = (10.0 / (100.0 / sqrt(96.0))) * 1.8;
        n_cand = 0;
        // HUD inputs (render_indicators_for_pygame :645-670)
        const double* wd = p.wheel + ((size_t)e * p.players + pi) * 8;
        hud_vals[0] = sqrt(vx * vx + vy * vy);
        for (int k = 0; k < 4; ++k) hud_vals[1 + k] = wd[k];
        hud_vals[5] = (double)(b[8 + 2] - b[2]);     // wheels[0].joint.angle
        hud_vals[6] = (double)b[5];                   // hull.angularVelocity
        hud_vals[7] = p.reward[2 * ((size_t)e * p.players + pi)];
    }
    __syncthreads();
    const Cam cm = cam;
    const int n_track = p.n_track[e];
    const CarTile* tiles = p.tiles + (size_t)e * CAR_MAX_TRACK;

    // ---- background: grass + checker squares at pixel centres.  fp64 only for the per-row terms; the
    //      per-pixel increment runs in fp32 on values bounded by +-24 (checker grid units), error ~1e-6 ----
    {
        const double kq = CR_PLAYFIELD / 20.0;
        const double inv = 1.0 / (cm.k * kq), ax = cm.camx / kq, ay = cm.camy / kq;
        const float cinv = (float)(cm.c_rot * inv), sinv = (float)(cm.s_rot * inv);
        for (int r = warp; r < CAR_H; r += RASTER_THREADS / 32) {
            const double dyp = r + 0.5 - CAR_H / 2.0;
            // g(c) = a - (dxp*c_rot - dyp*s_rot)*inv  with dxp = c + 0.5 - 48
            const double gx0d = ax + dyp * cm.s_rot * inv + (CAR_W / 2.0 - 0.5) * cm.c_rot * inv;
            const double gy0d = ay - dyp * cm.c_rot * inv + (CAR_W / 2.0 - 0.5) * cm.s_rot * inv;
            // keep magnitudes small for fp32: subtract the integer part of the row origin
            const double bx = floor(gx0d), by = floor(gy0d);
            const float fx0 = (float)(gx0d - bx), fy0 = (float)(gy0d - by);
            const int ibx = (int)fmax(fmin(bx, 1e6), -1e6), iby = (int)fmax(fmin(by, 1e6), -1e6);
            for (int c = lane; c < CAR_W; c += 32) {
                const int gx = ibx + (int)floorf(fx0 - (float)c * cinv), gy = iby + (int)floorf(fy0 - (float)c * sinv);
                const bool chk = gx >= -20 && gx < 20 && gy >= -20 && gy < 20 && !(gx & 1) && !(gy & 1);
                keys[r * KEY_STRIDE + c] = chk ? G[G_CHECK] : G[G_GRASS];
            }
        }
    }
    // ---- cull: tiles whose track point lies within the window's circumscribed circle (ordered by index) ----
    if (warp == 0) {
        const double reach = 48.0 * 1.4142135623730951 / cm.k + 2.0 * CR_TRACK_WIDTH + CR_BORDER + CR_TRACK_DETAIL_STEP;
        const float r2 = (float)(reach * reach);
        int base = 0;
        for (int t0 = 0; t0 < n_track; t0 += 32) {
            const int t = t0 + lane;
            bool in = false;
            if (t < n_track) {
                const float dx = tiles[t].cx - (float)cm.camx, dy = tiles[t].cy - (float)cm.camy;
                in = dx * dx + dy * dy <= r2;
            }
            const unsigned m = __ballot_sync(0xffffffffu, in);
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            if (in && pos < MAX_CAND) cand[pos] = t;
            base += __popc(m);
        }
        if (lane == 0) n_cand = min(base, MAX_CAND);
    }
    __syncthreads();
    // ---- road: paint order is tile n-1 .. 0, each followed by its kerb (:399-445); higher key wins ----
    {
        const int nc = n_cand;
        for (int q = warp; q < nc; q += RASTER_THREADS / 32) {
            const CarTile T = tiles[cand[q]];
            const int t = cand[q];
            const unsigned int order = 2u * (unsigned)(n_track - 1 - t) + 1u;
            float sx[8], sy[8];
            for (int i = 0; i < T.n; ++i) to_screen(cm, (double)T.px[i], (double)T.py[i], sx[i], sy[i]);
            const uint8_t g = G[G_ROAD0 + t % 3];
            fill_poly(keys, sx, sy, T.n, (order << 8) | g, lane);
            if (T.flags & 2) {
                for (int i = 0; i < 4; ++i) to_screen(cm, (double)T.kx[i], (double)T.ky[i], sx[i], sy[i]);
                fill_poly(keys, sx, sy, 4, ((order + 1u) << 8) | ((T.flags & 4) ? G[G_KERB_W] : G[G_KERB_R]), lane);
            }
        }
    }
    __syncthreads();
    // ---- cars: for k in cars: wheels, then hull fixtures (drawlist = wheels + [hull]) ----
    {
        const int per_car = 8;   // 4 wheels + 4 hull fixtures
        for (int q = warp; q < p.players * per_car; q += RASTER_THREADS / 32) {
            const int ck = q / per_car, part = q % per_car;
            const float* b = car_body[ck];
            const unsigned int order = 2048u + (unsigned)(ck * per_car + part);
            float sx[8], sy[8];
            int n;
            uint8_t g;
            if (part < 4) {
                const float* w = b + 8 * (part + 1);
                float ws, wc;
                sincosf(w[2], &ws, &wc);
                const float hw = (float)(14 * CR_SIZE), hr = (float)(27 * CR_SIZE);
                const float lx[4] = {-hw, +hw, +hw, -hw}, ly[4] = {-hr, -hr, +hr, +hr};
                for (int i = 0; i < 4; ++i)
                    to_screen(cm, (double)(wc * lx[i] - ws * ly[i] + w[0]), (double)(ws * lx[i] + wc * ly[i] + w[1]), sx[i], sy[i]);
                n = 4;
                g = G[G_WHEEL];
            } else {
                const int f = part - 4;
                float hs, hc;
                sincosf(b[2], &hs, &hc);
                const float hx = b[0] - (hc * K->hull_lcx - hs * K->hull_lcy), hy = b[1] - (hs * K->hull_lcx + hc * K->hull_lcy);
                n = c_hull_count[f];
                for (int i = 0; i < n; ++i) {
                    const float lx = (float)(c_hull_poly[f][i][0] * CR_SIZE), ly = (float)(c_hull_poly[f][i][1] * CR_SIZE);
                    to_screen(cm, (double)(hc * lx - hs * ly + hx), (double)(hs * lx + hc * ly + hy), sx[i], sy[i]);
                }
                g = (ck == pi) ? G[G_OWN] : G[G_OTHER];
            }
            fill_poly(keys, sx, sy, n, (order << 8) | g, lane);
        }
    }
    __syncthreads();
    for (int r = warp; r < CAR_H; r += RASTER_THREADS / 32)
        for (int c = lane; c < CAR_W; c += 32) img[r * CAR_W + c] = (uint8_t)(keys[r * KEY_STRIDE + c] & 255u);
    __syncthreads();
    // ---- HUD (painted after the scene) ----
    {
        const double W = CAR_W, H = CAR_H, s = W / 40.0, h = H / 40.0;
        hud_rect(img, 0, H - 4 * h, W, 4 * h * 1000, G[G_HUD], tid);
        __syncthreads();
        hud_rect(img, 5 * s, H - h, s, h * (-0.02 * hud_vals[0]), G[G_BLUE], tid);
        __syncthreads();
        for (int k = 0; k < 4; ++k) {
            hud_rect(img, (7 + k) * s, H - h, s, h * (-0.01 * hud_vals[1 + k]), k < 2 ? G[G_BLUE] : G[G_BLUE2], tid);
            __syncthreads();
        }
        hud_rect(img, 20 * s, H - 2 * h, s * (10.0 * hud_vals[5]), 2 * h, G[G_GREEN], tid);
        __syncthreads();
        hud_rect(img, 30 * s, H - 2 * h, s * (0.8 * hud_vals[6]), 2 * h, G[G_RED], tid);
        __syncthreads();
        if (tid == 0 && p.glyphs != nullptr) {     // draw_text("%05.0f" % reward) at (W/100, H - H/20)
            // "%05.0f": round half to even, sign kept for negative values, zero padded to width 5
            const double rv = hud_vals[7];
            double mag = rint(fabs(rv));
            char digits[24];
            int nd = 0;
            if (mag == 0) digits[nd++] = 0;
            while (mag >= 1 && nd < 20) { const double qd = floor(mag / 10.0); digits[nd++] = (char)(mag - qd * 10.0); mag = qd; }
            const int body = nd + (rv < 0 ? 1 : 0);
            int pen = (int)(W / 100);
            const int y0 = (int)(H - H / 20);
            const int pad = 5 > body ? 5 - body : 0;
            for (int i = 0; i < body + pad; ++i) {
                int gi;
                if (rv < 0 && i == 0) gi = 10;
                else if (i < (rv < 0 ? 1 : 0) + pad) gi = 0;
                else gi = digits[nd - 1 - (i - (rv < 0 ? 1 : 0) - pad)];
                for (int gy = 0; gy < 8; ++gy)
                    for (int gx = 0; gx < 4; ++gx)
                        if (p.glyphs[(gi * 8 + gy) * 4 + gx]) {
                            const int px = pen + gx, py = y0 + gy;
                            if (px >= 0 && px < CAR_W && py >= 0 && py < CAR_H) img[py * CAR_W + px] = G[G_TEXT];
                        }
                pen += p.glyphs[11 * 8 * 4 + gi];
            }
        }
        __syncthreads();
    }
    // ---- FrameStack: the new frame enters the ring; the observation is the ring oldest -> newest.
    //      After a reset (only_done pass, or the very first render) every slot holds the reset frame. ----
    const int C = p.c;
    uint8_t* ring = p.ring + ((size_t)e * p.players + pi) * C * CAR_PIX;
    const bool fill_all = only_done != 0 || p.ring_pos[e] < 0;
    const int newest = fill_all ? C - 1 : (p.ring_pos[e] + 1) % C;
    const uint4* src = reinterpret_cast<const uint4*>(img);
    if (fill_all) {
        for (int sl = 0; sl < C; ++sl) {
            uint4* dst = reinterpret_cast<uint4*>(ring + (size_t)sl * CAR_PIX);
            for (int q = tid; q < CAR_PIX / 16; q += RASTER_THREADS) dst[q] = src[q];
        }
    } else {
        uint4* dst = reinterpret_cast<uint4*>(ring + (size_t)newest * CAR_PIX);
        for (int q = tid; q < CAR_PIX / 16; q += RASTER_THREADS) dst[q] = src[q];
    }
    __syncthreads();
    // output layout: [env][players * C][96][96]: player-major channel blocks (FlattenMultiAgentObservation
    // concatenates the players' stacks on the channel axis), oldest frame first within a player
    uint8_t* out = obs + ((size_t)e * p.players + pi) * C * CAR_PIX;
    uint8_t* tout = (term_obs != nullptr && !only_done && p.env_done[e]) ? term_obs + ((size_t)e * p.players + pi) * C * CAR_PIX : nullptr;
    for (int sl = 0; sl < C; ++sl) {
        const int rs = fill_all ? sl : (newest + 1 + sl) % C;     // oldest first
        const uint4* rsrc = (rs == newest || fill_all) ? src : reinterpret_cast<const uint4*>(ring + (size_t)rs * CAR_PIX);
        uint4* dst = reinterpret_cast<uint4*>(out + (size_t)sl * CAR_PIX);
        uint4* tdst = tout ? reinterpret_cast<uint4*>(tout + (size_t)sl * CAR_PIX) : nullptr;
        for (int q = tid; q < CAR_PIX / 16; q += RASTER_THREADS) {
            const uint4 vv = rsrc[q];
            dst[q] = vv;
            if (tdst) tdst[q] = vv;
        }
    }
}

__global__ void car_ring_advance_kernel(CarDev p, int only_done) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n) return;
    if (only_done) {
        if (p.env_done[e]) p.ring_pos[e] = p.c - 1;
    } else {
        p.ring_pos[e] = p.ring_pos[e] < 0 ? p.c - 1 : (p.ring_pos[e] + 1) % p.c;
    }
}

cudaError_t launch_car_render(const CarDev& p, int only_done, uint8_t* obs, uint8_t* term_obs, cudaStream_t s) {
    car_render_kernel<<<p.n * p.players, RASTER_THREADS, 0, s>>>(p, only_done, obs, term_obs);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    car_ring_advance_kernel<<<(p.n + 127) / 128, 128, 0, s>>>(p, only_done);
    return cudaGetLastError();
}

}  // namespace crl
