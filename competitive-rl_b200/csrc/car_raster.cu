// car_raster.cu -- 96x96 grayscale observation of cCarRacing, one CTA per (env, player) frame.
//
// Replaces (paths relative to /root/reference/competitive_rl/): get_observation (car_racing_multi_players.py
// :622-634), camera_update (:791-812), camera_view (:764-789), render_road_for_observation_map (:732-755),
// render(mode="internal_rgb_array") (:857-863), Car.draw_for_pygame (car_dynamics.py:284-298),
// render_indicators_for_pygame (:645-670) + pygame_rendering.py:8-18, and the FrameStack /
// MultipleFrameStack + FlattenMultiAgentObservation + WrapPyTorch layout (utils/atari_wrappers.py:222-334).
//
// The reference paints the road once per reset into a 10 000 x 10 000 px surface (400 MB) at obs_scale
// px/unit, crops 192x192 around the camera, rotates the crop (pygame.transform.rotate, 16.16 fixed-point
// nearest neighbour) and centre-blits it to the 96x96 screen, then draws the cars' fixture polygons and the
// HUD.  No map is ever built here: every destination pixel is mapped through the same integer pipeline
// (screen -> rotated surface -> source crop -> road-map pixel (U, V)) and coloured by the last polygon in
// paint order whose pygame scanline fill covers (U, V).  Per polygon one warp builds the scanline span table
// (pygame 1.9 draw_fillpoly: C integer division per edge) in shared memory and sweeps the polygon's screen
// bounding box; atomicMax on (paint order << 8 | gray) lets all polygons fill in parallel while the
// reference's painter's order decides each pixel.  pygame's integer rules are restated from memory exactly
// as in oracle/ref_shim/pygame, under which the reference's own renderer reproduces these frames bit for bit
// (tests/golden/car_frames.npz); parity against a real pygame build is unpinned.
#include <math.h>

#include "car_common.cuh"

namespace crl {

#define CR_SIZE 0.02
#define CR_PLAYFIELD (2000.0 / 6.0)
#define CR_TRACK_WIDTH (40.0 / 6.0)
#define CR_BORDER (8.0 / 6.0)
#define CR_TRACK_DETAIL_STEP (21.0 / 6.0)

constexpr int RASTER_THREADS = 256;
constexpr int RASTER_WARPS = RASTER_THREADS / 32;
constexpr int MAX_CAND = 192;
constexpr int KEY_STRIDE = CAR_W + 1;      // padded: lanes of a pass touch different rows of the same columns
constexpr int SPAN_ROWS = 64;              // scanlines of one polygon (road tiles need <= ~30)

__constant__ float c_hull_poly[4][8][2] = {
    {{-60, +130}, {+60, +130}, {+60, +110}, {-60, +110}},
    {{-15, +120}, {+15, +120}, {+20, +20}, {-20, 20}},
    {{+25, +20}, {+50, -10}, {+50, -40}, {+20, -90}, {-20, -90}, {-50, -40}, {-50, -10}, {-25, +20}},
    {{-50, -120}, {+50, -120}, {+50, -90}, {-50, -90}}};
__constant__ int c_hull_count[4] = {4, 4, 8, 4};

// Integer parameters of one frame's screen -> road-map mapping (pygame.transform.rotate + blit + subsurface)
struct FrameMap {
    int rx, ry;            // top-left of the 192x192 crop in the road map
    int nx, ny;            // size of the rotated surface
    int bx, by;            // blit position of the rotated surface on the screen
    int cyi, isin, icos;
    int cx0, cy0;          // dx = cx0 + icos*x - isin*y ; dy = cy0 + isin*x + icos*y   (16.16)
    float inv_det;         // 1 / (icos^2 + isin^2), for the inverse mapping used to bound sweeps
    float camx, camy;      // camera_offset (b2Vec2)
    float ts, tc;          // sin/cos of tmp.angle = -camera_angle (fp32, b2Rot)
    double obs_scale;
};

// road-map pixel under screen pixel (X, Y); false: outside the rotated surface / source crop (black)
__device__ __forceinline__ bool map_pixel(const FrameMap& m, int X, int Y, int& U, int& V) {
    const int x = X - m.bx, y = Y - m.by;
    if (x < 0 || y < 0 || x >= m.nx || y >= m.ny) return false;
    const int dx = m.cx0 + m.icos * x - m.isin * y, dy = m.cy0 + m.isin * x + m.icos * y;
    if (dx < 0 || dy < 0 || dx > (2 * CAR_W << 16) - 1 || dy > (2 * CAR_H << 16) - 1) return false;
    U = m.rx + (dx >> 16);
    V = m.ry + (dy >> 16);
    return true;
}

// Same without the bounds tests: the visible 96x96 window is the centre of the rotated 192x192 crop, whose
// inscribed circle (radius 96) always contains it (half diagonal 68), so every screen pixel has a source.
__device__ __forceinline__ void map_pixel_nocheck(const FrameMap& m, int X, int Y, int& U, int& V) {
    const int x = X - m.bx, y = Y - m.by;
    U = m.rx + ((m.cx0 + m.icos * x - m.isin * y) >> 16);
    V = m.ry + ((m.cy0 + m.isin * x + m.icos * y) >> 16);
}

// approximate screen position of road-map point (u, v), to bound sweeps
__device__ __forceinline__ void map_to_screen(const FrameMap& m, float u, float v, float& X, float& Y) {
    const float dx = (u - (float)m.rx) * 65536.f - (float)m.cx0, dy = (v - (float)m.ry) * 65536.f - (float)m.cy0;
    X = ((float)m.icos * dx + (float)m.isin * dy) * m.inv_det + (float)m.bx;
    Y = (-(float)m.isin * dx + (float)m.icos * dy) * m.inv_det + (float)m.by;
}

// pygame 1.9 draw_fillpoly, one scanline: x spans (inclusive) of polygon (vx, vy)[n] at row V.
// Up to two spans (outlines with <= 8 vertices used here never give more); empty span = (1, 0).
__device__ __forceinline__ int4 scanline_spans(const int* vx, const int* vy, int n, int V, int maxy) {
    int xs[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    int m = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < n) {
            const int i1 = i ? i - 1 : n - 1;
            int y1 = vy[i1], y2 = vy[i], x1 = vx[i1], x2 = vx[i];
            if (y1 > y2) { int t = y1; y1 = y2; y2 = t; t = x1; x1 = x2; x2 = t; }
            if (y1 != y2 && ((V >= y1 && V < y2) || (V == maxy && V > y1 && V <= y2))) {
                const int x = (V - y1) * (x2 - x1) / (y2 - y1) + x1;      // C integer division
                if (m == 0) xs[0] = x; else if (m == 1) xs[1] = x; else if (m == 2) xs[2] = x; else if (m == 3) xs[3] = x;
                ++m;
            }
        }
    }
    // sort (unused slots hold INT_MAX): 4-element network
#define CSWAP(a, b) { const int lo_ = min(xs[a], xs[b]), hi_ = max(xs[a], xs[b]); xs[a] = lo_; xs[b] = hi_; }
    CSWAP(0, 1) CSWAP(2, 3) CSWAP(0, 2) CSWAP(1, 3) CSWAP(1, 2)
#undef CSWAP
    m = min(m, 4);
    int4 r = make_int4(1, 0, 1, 0);
    if (m >= 2) { r.x = xs[0]; r.y = xs[1]; }
    if (m >= 4) { r.z = xs[2]; r.w = xs[3]; }
    return r;
}

// Fill polygon (vx, vy)[n] given in the coordinate system `MAPPED ? road map : screen` into the key buffer.
template <bool MAPPED>
__device__ void fill_ipoly(unsigned int* keys, int4* spans, const FrameMap& fm, const int* vx, const int* vy, int n,
                           unsigned int key, int lane) {
    int minx = vx[0], maxx = vx[0], miny = vy[0], maxy = vy[0];
#pragma unroll
    for (int i = 1; i < 8; ++i)
        if (i < n) { minx = min(minx, vx[i]); maxx = max(maxx, vx[i]); miny = min(miny, vy[i]); maxy = max(maxy, vy[i]); }
    const int rows = maxy - miny + 1;
    if (rows > SPAN_ROWS || rows <= 0) return;            // not reachable for this geometry
    for (int r = lane; r < rows; r += 32) spans[r] = scanline_spans(vx, vy, n, miny + r, maxy);
    __syncwarp();
    int X0, X1, Y0, Y1;
    if (MAPPED) {
        // Every covered map pixel lies within one pixel of the polygon's hull, so the screen hull of the
        // (mapped) vertices, grown by the truncations along the way, bounds the sweep far tighter than the
        // rotated map-space box would.
        float fx0 = 1e9f, fx1 = -1e9f, fy0 = 1e9f, fy1 = -1e9f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < n) {
                float X, Y;
                map_to_screen(fm, (float)vx[i] + 0.5f, (float)vy[i] + 0.5f, X, Y);
                fx0 = fminf(fx0, X); fx1 = fmaxf(fx1, X); fy0 = fminf(fy0, Y); fy1 = fmaxf(fy1, Y);
            }
        }
        X0 = max(0, (int)floorf(fx0 - 2.5f)); X1 = min(CAR_W - 1, (int)ceilf(fx1 + 2.5f));
        Y0 = max(0, (int)floorf(fy0 - 2.5f)); Y1 = min(CAR_H - 1, (int)ceilf(fy1 + 2.5f));
    } else {
        X0 = max(0, minx); X1 = min(CAR_W - 1, maxx); Y0 = max(0, miny); Y1 = min(CAR_H - 1, maxy);
    }
    if (X1 >= X0 && Y1 >= Y0) {
        const int bw = X1 - X0 + 1;
        const int lw = (bw <= 8) ? 3 : (bw <= 16) ? 4 : 5;          // pass shape 4x8, 2x16 or 1x32
        const int lx = lane & ((1 << lw) - 1), ly = lane >> lw, rows_per_pass = 32 >> lw;
        for (int yb = Y0; yb <= Y1; yb += rows_per_pass) {
            const int Y = yb + ly;
            for (int xb = X0; xb <= X1; xb += (1 << lw)) {
                const int X = xb + lx;
                int U = X, V = Y;
                bool ok = (Y <= Y1) && (X <= X1);
                if (MAPPED) map_pixel_nocheck(fm, X, Y, U, V);
                const int r = V - miny;
                if (ok && r >= 0 && r < rows) {
                    const int4 sp = spans[r];
                    if ((U >= sp.x && U <= sp.y) || (U >= sp.z && U <= sp.w)) atomicMax(&keys[Y * KEY_STRIDE + X], key);
                }
            }
        }
    }
    __syncwarp();
}

// pygame.draw.rect(screen, color, (x, y, w, h)) = polygon (l, t), (r, t), (r, b), (l, b), r = x + w - 1, b = y + h - 1
__device__ __forceinline__ void hud_rect(uint8_t* img, double x, double y, double w, double h, uint8_t val, int tid) {
    const int X = (int)x, Y = (int)y, W = (int)w, H = (int)h;
    const int l = X, r = X + W - 1, t = Y, b = Y + H - 1;
    if (t == b) return;                                   // every edge horizontal or degenerate: nothing is filled
    const int x0 = max(min(l, r), 0), x1 = min(max(l, r), CAR_W - 1), y0 = max(min(t, b), 0), y1 = min(max(t, b), CAR_H - 1);
    const int bw = x1 - x0 + 1, total = bw * (y1 - y0 + 1);
    if (bw <= 0 || total <= 0) return;
    for (int q = tid; q < total; q += RASTER_THREADS) img[(y0 + q / bw) * CAR_W + x0 + q % bw] = val;
}

struct RasterSmem {
    unsigned int keys[CAR_H * KEY_STRIDE];
    uint8_t img[CAR_PIX];
    int4 spans[RASTER_WARPS][SPAN_ROWS];
    int cand[MAX_CAND];
    uint8_t chk_x[2 * CAR_W], chk_y[2 * CAR_H];   // is road-map column rx + i / row ry + i inside a checker square
    float car_body[CAR_MAX_PLAYERS][40];
    double hud_vals[8];
    FrameMap fm;
    int n_cand;
};

__global__ void __launch_bounds__(RASTER_THREADS)
car_render_kernel(CarDev p, int only_done, uint8_t* __restrict__ obs, uint8_t* __restrict__ term_obs) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    RasterSmem& S = *reinterpret_cast<RasterSmem*>(smem_raw);
    const int frame = blockIdx.x;                     // env * players + player
    const int e = frame / p.players, pi = frame % p.players;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (only_done && !p.env_done[e]) return;
    const CarHullConst* K = p.consts;
    const uint8_t* G = K->gray;
    const int* checker = K->checker;

    if (tid < p.players * 40) S.car_body[tid / 40][tid % 40] = p.body[((size_t)e * p.players + tid / 40) * 40 + tid % 40];
    __syncthreads();
    if (tid == 0) {
        // ---- camera_update("rgb_array"): hull.position + R(angle) * (0, 16) ----
        const float* b = S.car_body[pi];
        float hs, hc;
        sincosf(b[2], &hs, &hc);
        const float hx = b[0] - (hc * K->hull_lcx - hs * K->hull_lcy), hy = b[1] - (hs * K->hull_lcx + hc * K->hull_lcy);
        double angle = (double)b[2];
        const double vx = (double)b[3], vy = (double)b[4];
        if (vx * vx + vy * vy > 0.5 * 0.5) angle = atan2(-vx, +vy);
        const float fa = (float)angle;
        float fs, fc;
        sincosf(fa, &fs, &fc);
        FrameMap m;
        m.obs_scale = (10 / (100 / sqrt(96.0))) * 1.8;
        m.camx = hx + (fc * 0.0f - fs * 16.0f);
        m.camy = hy + (fs * 0.0f + fc * 16.0f);
        // ---- camera_view: crop rectangle, rotation, blit ----
        const double pos0 = m.obs_scale * -(double)m.camx + 5000.0, pos1 = m.obs_scale * -(double)m.camy + 5000.0;
        m.rx = (int)(pos0 - CAR_W); m.ry = (int)(pos1 - CAR_H);
        const int sw = 2 * CAR_W, sh = 2 * CAR_H;
        const double rad = (57.295779513 * angle) * .01745329251994329;
        const double sangle = sin(rad), cangle = cos(rad);
        const double cx = cangle * sw, cy = cangle * sh, sx = sangle * sw, sy = sangle * sh;
        m.nx = (int)fmax(fmax(fmax(fabs(cx + sy), fabs(cx - sy)), fabs(-cx + sy)), fabs(-cx - sy));
        m.ny = (int)fmax(fmax(fmax(fabs(sx + cy), fabs(sx - cy)), fabs(-sx + cy)), fabs(-sx - cy));
        m.cyi = m.ny / 2;
        const int xd = (sw - m.nx) * 32768, yd = (sh - m.ny) * 32768;
        m.isin = (int)(sangle * 65536); m.icos = (int)(cangle * 65536);
        const int ax = (m.nx * 32768) - (int)(cangle * (double)((m.nx - 1) * 32768));
        const int ay = (m.ny * 32768) - (int)(sangle * (double)((m.nx - 1) * 32768));
        m.cx0 = ax + xd + m.isin * m.cyi;
        m.cy0 = ay + yd - m.icos * m.cyi;
        m.bx = -(m.nx >> 1) + CAR_W / 2; m.by = -(m.ny >> 1) + CAR_H / 2;
        m.inv_det = 1.0f / ((float)m.icos * (float)m.icos + (float)m.isin * (float)m.isin);
        const float ta = (float)(-angle);
        sincosf(ta, &m.ts, &m.tc);
        S.fm = m;
        S.n_cand = 0;
        // HUD inputs (render_indicators_for_pygame :645-670)
        const double* wd = p.wheel + ((size_t)e * p.players + pi) * 8;
        S.hud_vals[0] = sqrt(vx * vx + vy * vy);
        for (int k = 0; k < 4; ++k) S.hud_vals[1 + k] = wd[k];
        S.hud_vals[5] = (double)(b[8 + 2] - b[2]);     // wheels[0].joint.angle
        S.hud_vals[6] = (double)b[5];                   // hull.angularVelocity
        S.hud_vals[7] = p.reward[2 * ((size_t)e * p.players + pi)];
    }
    __syncthreads();
    const FrameMap fm = S.fm;
    const int n_track = p.n_track[e];
    const CarTile* tiles = p.tiles + (size_t)e * CAR_MAX_TRACK;
    unsigned int* keys = S.keys;

    // ---- background: black outside the source, else grass / checker squares (:733-746) by road-map pixel.
    //      The squares are axis-aligned in the road map: tabulate per crop column / row whether it lies in one. ----
    for (int i = tid; i < 2 * CAR_W + 2 * CAR_H; i += RASTER_THREADS) {
        const bool is_y = i >= 2 * CAR_W;
        const int v = is_y ? fm.ry + (i - 2 * CAR_W) : fm.rx + i;
        const int* tab = checker + (is_y ? 40 : 0);
        bool in = false;
#pragma unroll 4
        for (int c = 0; c < 20; ++c) in = in || (v >= tab[2 * c] && v <= tab[2 * c + 1]);
        if (is_y) S.chk_y[i - 2 * CAR_W] = in; else S.chk_x[i] = in;
    }
    __syncthreads();
    for (int Y = warp; Y < CAR_H; Y += RASTER_WARPS) {
        for (int X = lane; X < CAR_W; X += 32) {
            int U, V;
            unsigned int key = 0u;    // surfaces start black
            if (map_pixel(fm, X, Y, U, V)) key = (S.chk_x[U - fm.rx] && S.chk_y[V - fm.ry]) ? G[G_CHECK] : G[G_GRASS];
            keys[Y * KEY_STRIDE + X] = key;
        }
    }
    // ---- cull: tiles that can reach the visible window (the central 96x96 of the rotated crop; ordered by index) ----
    if (warp == 0) {
        const double reach = (48.0 * 1.4142135623730951 + 4.0) / fm.obs_scale + 2.0 * CR_TRACK_WIDTH + CR_BORDER + CR_TRACK_DETAIL_STEP;
        const float r2 = (float)(reach * reach);
        int base = 0;
        for (int t0 = 0; t0 < n_track; t0 += 32) {
            const int t = t0 + lane;
            bool in = false;
            if (t < n_track) {
                const float dx = tiles[t].cx - fm.camx, dy = tiles[t].cy - fm.camy;
                in = dx * dx + dy * dy <= r2;
            }
            const unsigned m = __ballot_sync(0xffffffffu, in);
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            if (in && pos < MAX_CAND) S.cand[pos] = t;
            base += __popc(m);
        }
        if (lane == 0) S.n_cand = min(base, MAX_CAND);
    }
    __syncthreads();
    // ---- road: paint order is tile n-1 .. 0, each followed by its kerb (:399-445); higher key wins ----
    {
        const int nc = S.n_cand;
        for (int q = warp; q < nc; q += RASTER_WARPS) {
            const int t = S.cand[q];
            const CarTile T = tiles[t];
            const unsigned int order = 2u * (unsigned)(n_track - 1 - t) + 1u;
            int vx[8], vy[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { vx[i] = (i < 5) ? T.mx[i] : 0; vy[i] = (i < 5) ? T.my[i] : 0; }
            fill_ipoly<true>(keys, S.spans[warp], fm, vx, vy, 5, (order << 8) | G[G_ROAD0 + t % 3], lane);
            if (T.flags & 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { vx[i] = T.kmx[i]; vy[i] = T.kmy[i]; }
                fill_ipoly<true>(keys, S.spans[warp], fm, vx, vy, 4, ((order + 1u) << 8) | ((T.flags & 4) ? G[G_KERB_W] : G[G_KERB_R]), lane);
            }
        }
    }
    __syncthreads();
    // ---- cars (Car.draw_for_pygame): for k in cars: wheels, then hull fixtures; b2Vec2 fp32 arithmetic:
    //      path = -scale * (tmp * ((trans * v) - offset)) + (W/2, H/2), truncated to int by pygame ----
    {
        const int per_car = 8;   // 4 wheels + 4 hull fixtures
        for (int q = warp; q < p.players * per_car; q += RASTER_WARPS) {
            const int ck = q / per_car, part = q % per_car;
            const float* b = S.car_body[ck];
            const unsigned int order = 2048u + (unsigned)(ck * per_car + part);
            const float* body = (part < 4) ? b + 8 * (part + 1) : b;
            float bs, bc;
            sincosf(body[2], &bs, &bc);
            float px = body[0], py = body[1];
            if (part >= 4) { px = b[0] - (bc * K->hull_lcx - bs * K->hull_lcy); py = b[1] - (bs * K->hull_lcx + bc * K->hull_lcy); }
            int vx[8], vy[8];
            const int n = (part < 4) ? 4 : c_hull_count[part - 4];
            const float hw = (float)(14 * CR_SIZE), hr = (float)(27 * CR_SIZE);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                vx[i] = 0; vy[i] = 0;
                if (i < n) {
                    float lx, ly;
                    if (part < 4) { lx = (i == 0 || i == 1) ? hw : -hw; ly = (i == 1 || i == 2) ? hr : -hr; }
                    else { lx = (float)(c_hull_poly[part - 4][i][0] * CR_SIZE); ly = (float)(c_hull_poly[part - 4][i][1] * CR_SIZE); }
                    const float wx = (bc * lx - bs * ly) + px, wy = (bs * lx + bc * ly) + py;
                    const float ox = wx - fm.camx, oy = wy - fm.camy;
                    const float rx2 = (fm.tc * ox - fm.ts * oy) + 0.0f, ry2 = (fm.ts * ox + fm.tc * oy) + 0.0f;
                    const float sxp = (float)((double)rx2 * -fm.obs_scale) + (float)(CAR_W / 2.0);
                    const float syp = (float)((double)ry2 * -fm.obs_scale) + (float)(CAR_H / 2.0);
                    vx[i] = (int)sxp; vy[i] = (int)syp;
                }
            }
            const uint8_t g = (part < 4) ? G[G_WHEEL] : ((ck == pi) ? G[G_OWN] : G[G_OTHER]);
            fill_ipoly<false>(keys, S.spans[warp], fm, vx, vy, n, (order << 8) | g, lane);
        }
    }
    __syncthreads();
    uint8_t* img = S.img;
    for (int r = warp; r < CAR_H; r += RASTER_WARPS)
        for (int c = lane; c < CAR_W; c += 32) img[r * CAR_W + c] = (uint8_t)(keys[r * KEY_STRIDE + c] & 255u);
    __syncthreads();
    // ---- HUD (painted after the scene) ----
    {
        const double W = CAR_W, H = CAR_H, s = W / 40.0, h = H / 40.0;
        hud_rect(img, 0, H - 4 * h, W, 4 * h * 1000, G[G_HUD], tid);
        __syncthreads();
        hud_rect(img, 5 * s, H - h, s, h * (-0.02 * S.hud_vals[0]), G[G_BLUE], tid);
        __syncthreads();
        for (int k = 0; k < 4; ++k) {
            hud_rect(img, (7 + k) * s, H - h, s, h * (-0.01 * S.hud_vals[1 + k]), k < 2 ? G[G_BLUE] : G[G_BLUE2], tid);
            __syncthreads();
        }
        hud_rect(img, 20 * s, H - 2 * h, s * (10.0 * S.hud_vals[5]), 2 * h, G[G_GREEN], tid);
        __syncthreads();
        hud_rect(img, 30 * s, H - 2 * h, s * (0.8 * S.hud_vals[6]), 2 * h, G[G_RED], tid);
        __syncthreads();
        if (tid == 0 && p.glyphs != nullptr) {     // draw_text("%05.0f" % reward) at (W/100, H - H/20)
            // "%05.0f": round half to even, sign kept for negative values, zero padded to width 5
            const double rv = S.hud_vals[7];
            double mag = rint(fabs(rv));
            char digits[24];
            int nd = 0;
            if (mag == 0) digits[nd++] = 0;
            while (mag >= 1 && nd < 20) { const double qd = floor(mag / 10.0); digits[nd++] = (char)(mag - qd * 10.0); mag = qd; }
            const int neg = (rv < 0 || (rv == 0 && signbit(rv))) ? 1 : 0;
            const int body = nd + neg;
            int pen = (int)(W / 100);
            const int y0 = (int)(H - H / 20);
            const int pad = 5 > body ? 5 - body : 0;
            for (int i = 0; i < body + pad; ++i) {
                int gi;
                if (neg && i == 0) gi = 10;
                else if (i < neg + pad) gi = 0;
                else gi = digits[nd - 1 - (i - neg - pad)];
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

// road-map bounds of the checker squares: for even grid cells g = -20 + 2i the square polygon
// (k*g + k, .), (k*g, .) ... is int-truncated by pygame; per axis [lo_i, hi_i] inclusive (host, fp64)
void car_checker_table(int* out /* [2][20][2] */) {
    const double k = CR_PLAYFIELD / 20.0, osc = (10 / (100 / sqrt(96.0))) * 1.8;
    for (int axis = 0; axis < 2; ++axis)
        for (int i = 0; i < 20; ++i) {
            const int g = -20 + 2 * i;
            const int a = (int)(osc * -(k * g + k) + 5000.0), b = (int)(osc * -(k * g + 0) + 5000.0);
            out[(axis * 20 + i) * 2] = a < b ? a : b;
            out[(axis * 20 + i) * 2 + 1] = a < b ? b : a;
        }
}

cudaError_t car_raster_init() {
    return cudaFuncSetAttribute(car_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RasterSmem));
}

cudaError_t launch_car_render(const CarDev& p, int only_done, uint8_t* obs, uint8_t* term_obs, cudaStream_t s) {
    car_render_kernel<<<p.n * p.players, RASTER_THREADS, sizeof(RasterSmem), s>>>(p, only_done, obs, term_obs);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    car_ring_advance_kernel<<<(p.n + 127) / 128, 128, 0, s>>>(p, only_done);
    return cudaGetLastError();
}

}  // namespace crl
