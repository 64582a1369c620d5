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
// paint order whose pygame scanline fill covers (U, V).  Every polygon that can reach the window (road tiles and
// kerbs in road-map pixels, car fixtures in screen pixels) gets its scanline span table (pygame 1.9
// draw_fillpoly: C integer division per edge) in one shared-memory pool and is binned, by the screen bounding
// box of its vertices, into 8x8-pixel cells; a warp then walks a cell with two pixels per lane and tests only
// that cell's polygons, keeping the largest (paint order << 8 | gray) -- the reference's painter's order --
// per pixel in registers.  pygame's integer rules are restated from memory exactly
// as in oracle/ref_shim/pygame, under which the reference's own renderer reproduces these frames bit for bit
// (tests/golden/car_frames.npz); parity against a real pygame build is unpinned.
#include <math.h>

#include "car_common.cuh"
#include "car_spans.cuh"

namespace crl {

#define CR_SIZE 0.02
#define CR_PLAYFIELD (2000.0 / 6.0)
#define CR_TRACK_WIDTH (40.0 / 6.0)
#define CR_BORDER (8.0 / 6.0)
#define CR_TRACK_DETAIL_STEP (21.0 / 6.0)

constexpr int RASTER_THREADS = 256;
constexpr int RASTER_WARPS = RASTER_THREADS / 32;
constexpr int MAX_POLY = 240;              // 224 road polygons (tiles + kerbs) + 16 car fixtures of one frame (ids fit a byte)
constexpr int POOL_ROWS = 2048;            // scanline span table shared by all polygons of a frame
constexpr int CELL = 8, CELLS_X = CAR_W / CELL;
constexpr int WALK_CELLS = CELLS_X * ((86 + CELL - 1) / CELL);   // cells with rows above the HUD bar (HUD_TOP = 86): 11 rows of 12
constexpr int MASK_WORDS = (MAX_POLY + 31) / 32;   // per-cell bitmask over the polygon ids
constexpr int CAR_WORD = MASK_WORDS - 1;   // ids CAR_WORD * 32 .. are the car fixtures (screen space); below: road (map space)
constexpr int MAX_ROAD_POLY = CAR_WORD * 32;
constexpr unsigned short NO_TABLE = 0xFFFFu;
constexpr int HUD_TOP = 86;                // (int)(H - 4 * (H / 40.0)) = (int)86.4: first row of the black HUD bar
constexpr int HUD_WARP = 5;                // the warp that paints the HUD (it has no other work unless a frame has > 160 road tiles)

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
};
__device__ __forceinline__ double car_obs_scale() { return (10 / (100 / sqrt(96.0))) * 1.8; }   // CarRacing.obs_scale, :215

// Screen pixel (X, Y) -> road-map pixel: with (x, y) = (X - bx, Y - by), dx = cx0 + icos * x - isin * y and
// dy = cy0 + isin * x + icos * y in 16.16 fixed point, (U, V) = (rx + (dx >> 16), ry + (dy >> 16)) -- evaluated
// incrementally in walk_cells.  The visible 96x96 window is the centre of the rotated 192x192 crop, whose inscribed
// circle (radius 96) always contains it (half diagonal 68), so every screen pixel has a source inside the crop.

// approximate screen position of road-map point (u, v), to bound sweeps
__device__ __forceinline__ void map_to_screen(const FrameMap& m, float u, float v, float& X, float& Y) {
    const float dx = (u - (float)m.rx) * 65536.f - (float)m.cx0, dy = (v - (float)m.ry) * 65536.f - (float)m.cy0;
    X = ((float)m.icos * dx + (float)m.isin * dy) * m.inv_det + (float)m.bx;
    Y = (-(float)m.isin * dx + (float)m.icos * dy) * m.inv_det + (float)m.by;
}

// pygame.draw.rect(screen, color, (x, y, w, h)) = polygon (l, t), (r, t), (r, b), (l, b), r = x + w - 1, b = y + h - 1
__device__ __forceinline__ void hud_rect(uint8_t* img, double x, double y, double w, double h, uint8_t val, int tid, int nthreads) {
    const int X = (int)x, Y = (int)y, W = (int)w, H = (int)h;
    const int l = X, r = X + W - 1, t = Y, b = Y + H - 1;
    if (t == b) return;                                   // every edge horizontal or degenerate: nothing is filled
    const int x0 = max(min(l, r), 0), x1 = min(max(l, r), CAR_W - 1), y0 = max(min(t, b), 0), y1 = min(max(t, b), CAR_H - 1);
    const int bw = x1 - x0 + 1, total = bw * (y1 - y0 + 1);
    if (bw <= 0 || total <= 0) return;
    for (int q = tid; q < total; q += nthreads) img[(y0 + q / bw) * CAR_W + x0 + q % bw] = val;
}

// one polygon of the frame: rows [miny, miny + rows) of its span table start at spans[off] (NO_TABLE: the pool was
// full, spans are recomputed per pixel); screen = 1: vertices are screen pixels (cars), else road-map pixels
struct __align__(16) PolyMeta { short miny, rows; unsigned short off; unsigned char n, screen; unsigned int key; unsigned int pad; };

struct RasterSmem {
    uint8_t img[CAR_PIX];
    short4 spans[POOL_ROWS];
    uint8_t row_owner[POOL_ROWS];
    short pvx[MAX_POLY - MAX_ROAD_POLY][8], pvy[MAX_POLY - MAX_ROAD_POLY][8];   // vertices of the car polygons (road: CarTile)
    const CarTile* tiles;                          // the env's tiles (road polygon vertices) and their span tables
    const short4* env_spans;
    PolyMeta meta[MAX_POLY];
    uint32_t cell_mask[WALK_CELLS][MASK_WORDS];      // polygons whose screen bounding box touches the cell
    uint8_t chk_x[2 * CAR_W], chk_y[2 * CAR_H];   // is road-map column rx + i / row ry + i inside a checker square
    float car_body[CAR_MAX_PLAYERS][40];
    double hud_vals[8];
    FrameMap fm;
    int n_poly, pool_used, overflow;
    int copy_next, hud_late;                       // ring -> observation chunk counter; 1 = an indicator reaches above the bar
};

// Register polygon (vx, vy)[n] of the frame: id, span-table rows, cell bins.  One thread per polygon.
// `cache` = 0x40000000 | kind << 16 | tile for a road polygon (its vertices live in CarTile; bit 31 is added here when its
// span table fits CarDev::tile_spans), 0 for a car polygon (vertices kept in shared memory).
__device__ void add_polygon(RasterSmem& S, const FrameMap& fm, const int* vx, const int* vy, int n, unsigned int key, bool screen,
                            int car_slot, unsigned int cache) {
    int minx = vx[0], maxx = vx[0], miny = vy[0], maxy = vy[0];
#pragma unroll
    for (int i = 1; i < 8; ++i)
        if (i < n) { minx = min(minx, vx[i]); maxx = max(maxx, vx[i]); miny = min(miny, vy[i]); maxy = max(maxy, vy[i]); }
    const int rows = maxy - miny + 1;
    int X0, X1, Y0, Y1;
    if (!screen) {
        // Every covered map pixel lies within one pixel of the polygon's hull, so the screen hull of the (mapped)
        // vertices, grown by the truncations along the way, bounds the covered screen pixels.
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
    if (X1 < X0 || Y1 < Y0 || rows <= 0) return;          // cannot touch the window
    const int id = screen ? MAX_ROAD_POLY + car_slot : atomicAdd(&S.n_poly, 1);
    if (!screen && id >= MAX_ROAD_POLY) { S.overflow = 2; return; }       // polygon dropped (reported through crl_car_check)
    int off = atomicAdd(&S.pool_used, rows);
    if (off + rows > POOL_ROWS) { off = NO_TABLE; S.overflow = 1; }
    else for (int r = 0; r < rows; ++r) S.row_owner[off + r] = (uint8_t)id;
    PolyMeta m;
    m.miny = (short)miny; m.rows = (short)rows; m.off = (unsigned short)off; m.n = (unsigned char)n; m.screen = screen ? 1 : 0; m.key = key;
    m.pad = cache | ((cache != 0u && rows <= (((cache >> 16) & 1u) ? CAR_SPAN_KERB_ROWS : CAR_SPAN_TILE_ROWS)) ? 0x80000000u : 0u);
    S.meta[id] = m;
    if (screen) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { S.pvx[car_slot][i] = (short)max(-32000, min(32000, vx[i])); S.pvy[car_slot][i] = (short)max(-32000, min(32000, vy[i])); }
    }
    const uint32_t bit = 1u << (id & 31);
    for (int cy = Y0 / CELL; cy <= min(Y1, HUD_TOP - 1) / CELL; ++cy)      // rows under the HUD bar are never walked
        for (int cx = X0 / CELL; cx <= X1 / CELL; ++cx) atomicOr(&S.cell_mask[cy * CELLS_X + cx][id >> 5], bit);
}

// spans of row V of polygon `id` without the pooled table: from the per-reset cache, or scanned from the vertices
__device__ __forceinline__ short4 poly_row_spans(const RasterSmem& S, int id, const PolyMeta& m, int V) {
    const int r = V - m.miny;
    if (m.pad & 0x80000000u)
        return S.env_spans[(size_t)(m.pad & 0xFFFFu) * CAR_SPAN_ROWS + (((m.pad >> 16) & 1u) ? CAR_SPAN_TILE_ROWS : 0) + r];
    if (m.pad & 0x40000000u) {
        const CarTile* T = S.tiles + (m.pad & 0xFFFFu);
        const bool kerb = (m.pad >> 16) & 1u;
        return scanline_spans(kerb ? T->kmx : T->mx, kerb ? T->kmy : T->my, m.n, V, m.miny + m.rows - 1);
    }
    return scanline_spans(S.pvx[id - MAX_ROAD_POLY], S.pvy[id - MAX_ROAD_POLY], m.n, V, m.miny + m.rows - 1);
}

// Pixels of the frame above the HUD bar: a warp walks one 8x8 cell at a time, lane = (x, y) and (x, y + 4).  Background: black outside
// the source crop, else grass / checker by road-map pixel; then the polygons binned to the cell, largest key wins.
// SLOW (the span pool overflowed): polygons without a table get their spans recomputed per pixel.
// one polygon against the two pixels of a lane; (xa, ya) / (xb, yb) in the polygon's coordinate system
template <bool SLOW>
__device__ __forceinline__ void test_polygon(const RasterSmem& S, int id, int xa, int ya, int xb, int yb, unsigned int& ka,
                                             unsigned int& kb) {
    const PolyMeta m = S.meta[id];
    if (m.key < ka && m.key < kb) return;
    const int ra = ya - m.miny, rb = yb - m.miny;
    if ((unsigned)ra < (unsigned)m.rows && m.key > ka) {
        const short4 sp = (!SLOW || m.off != NO_TABLE) ? S.spans[m.off + ra] : poly_row_spans(S, id, m, ya);
        if ((xa >= sp.x && xa <= sp.y) || (xa >= sp.z && xa <= sp.w)) ka = m.key;
    }
    if ((unsigned)rb < (unsigned)m.rows && m.key > kb) {
        const short4 sp = (!SLOW || m.off != NO_TABLE) ? S.spans[m.off + rb] : poly_row_spans(S, id, m, yb);
        if ((xb >= sp.x && xb <= sp.y) || (xb >= sp.z && xb <= sp.w)) kb = m.key;
    }
}

template <bool SLOW>
__device__ __forceinline__ void walk_cells(RasterSmem& S, const FrameMap& fm, unsigned int g_grass, unsigned int g_check,
                                           int warp, int lane) {
    uint8_t* img = S.img;
    const int n_words = (min(S.n_poly, MAX_ROAD_POLY) + 31) >> 5;
    const int lx = lane & 7, ly = lane >> 3;
    // dx = cx0 + icos * x - isin * y, dy = cy0 + isin * x + icos * y with (x, y) = (X - bx, Y - by): affine in the lane and
    // in the cell.  No bounds test: the 96x96 window is the centre of the rotated 192x192 crop, whose inscribed circle
    // (radius 96) contains it (half diagonal 68), so every screen pixel samples inside the crop.
    const int ldx = fm.cx0 + fm.icos * (lx - fm.bx) - fm.isin * (ly - fm.by);
    const int ldy = fm.cy0 + fm.isin * (lx - fm.bx) + fm.icos * (ly - fm.by);
    // cells warp, warp + 8, ...: 12 cells per row of cells, so +8 cells = +64 px in x, wrapping into the next row
    int cx = warp * CELL, cy = 0;
    for (int cell = warp; cell < WALK_CELLS; cell += RASTER_WARPS, cx += RASTER_WARPS * CELL) {
        if (cx >= CAR_W) { cx -= CAR_W; cy += CELL; }
        const int X = cx + lx, Ya = cy + ly, Yb = Ya + 4;
        uint8_t* pa = img + Ya * CAR_W + X;
        const int dxa = ldx + fm.icos * cx - fm.isin * cy, dya = ldy + fm.isin * cx + fm.icos * cy;
        const int dxb = dxa - 4 * fm.isin, dyb = dya + 4 * fm.icos;
        const int ua = (dxa >> 16) & 255, va = (dya >> 16) & 255, ub = (dxb >> 16) & 255, vb = (dyb >> 16) & 255;   // 0..191 (see above)
        const int Ua = fm.rx + ua, Va = fm.ry + va, Ub = fm.rx + ub, Vb = fm.ry + vb;
        unsigned int ka = (S.chk_x[ua] & S.chk_y[va]) ? g_check : g_grass;
        unsigned int kb = (S.chk_x[ub] & S.chk_y[vb]) ? g_check : g_grass;
        for (int w = 0; w < n_words; ++w) {                 // road tiles and kerbs: road-map coordinates
            unsigned int bits = S.cell_mask[cell][w];
            while (bits) {
                const int id = w * 32 + __ffs(bits) - 1;
                bits &= bits - 1u;
                test_polygon<SLOW>(S, id, Ua, Va, Ub, Vb, ka, kb);
            }
        }
        {                                                   // car fixtures: screen coordinates
            unsigned int bits = S.cell_mask[cell][CAR_WORD];
            while (bits) {
                const int id = CAR_WORD * 32 + __ffs(bits) - 1;
                bits &= bits - 1u;
                test_polygon<SLOW>(S, id, X, Ya, X, Yb, ka, kb);
            }
        }
        pa[0] = (uint8_t)(ka & 255u);                        // Ya <= 83
        if (Yb < HUD_TOP) pa[4 * CAR_W] = (uint8_t)(kb & 255u);   // rows 86.. belong to the HUD bar (painted by the HUD warp)
    }
}

// Per-frame setup, one warp per (env, player) frame: camera and the integer screen -> road-map mapping (lane 0), then
// the cull of the road tiles against the visible window (all lanes).  Kept out of the render kernel, where these
// serial steps would stall a whole CTA.
__global__ void __launch_bounds__(RASTER_THREADS)
car_frame_setup_kernel(CarDev p, int only_done, int which) {
    __shared__ FrameMap s_fm[RASTER_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int frame = blockIdx.x * RASTER_WARPS + warp;            // env * players + player
    if (frame >= p.n * p.players) return;
    const int e = frame / p.players;
    if (only_done && !p.env_done[e]) return;
    if (which != 0 && (p.deferred[e] != 0) != (which == 2)) return;
    const CarHullConst* K = p.consts;
    if (lane == 0) {
        // ---- camera_update("rgb_array"): hull.position + R(angle) * (0, 16) ----
        const float* b = p.body + (size_t)frame * 40;
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
        const double obs_scale = car_obs_scale();
        m.camx = hx + (fc * 0.0f - fs * 16.0f);
        m.camy = hy + (fs * 0.0f + fc * 16.0f);
        // ---- camera_view: crop rectangle, rotation, blit ----
        const double pos0 = obs_scale * -(double)m.camx + 5000.0, pos1 = obs_scale * -(double)m.camy + 5000.0;
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
        s_fm[warp] = m;
        p.frame_map[frame] = m;
    }
    __syncwarp();
    const FrameMap fm = s_fm[warp];
    // ---- cull: road tiles whose centre, mapped to the screen, lies within the window grown by the tile's reach
    //      (farthest kerb corner 8.7 units = 15.4 px, plus the slack of the integer pipeline) ----
    const int slot = car_slot(p, e);
    const int n_track = p.n_track[slot];
    const float2* centres = p.tile_centres + (size_t)slot * CAR_MAX_TRACK;
    uint16_t* cand = p.frame_cand + (size_t)frame * CAR_MAX_CAND;
    const double obs_scale = car_obs_scale();
    const float reach = 20.0f;
    int base = 0;
    for (int t0 = 0; t0 < n_track; t0 += 32) {
        const int t = t0 + lane;
        bool in = false;
        if (t < n_track) {
            const float2 tc = centres[t];                 // coalesced (CarTile is 116 bytes: one sector per lane otherwise)
            const float u = (float)(obs_scale * -(double)tc.x + 5000.0), v = (float)(obs_scale * -(double)tc.y + 5000.0);
            float X, Y;
            map_to_screen(fm, u, v, X, Y);
            in = X > -reach && X < CAR_W + reach && Y > -reach && Y < CAR_H + reach;
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        if (in && pos < CAR_MAX_CAND) cand[pos] = (uint16_t)t;
        base += __popc(m);
    }
    if (lane == 0) {
        p.frame_ncand[frame] = min(base, CAR_MAX_CAND);
        if (base > CAR_MAX_CAND) atomicAdd(p.overrun + 1, 1);   // tiles dropped: crl_car_check reports it
    }
}

// HUD indicators and reward text (render_indicators_for_pygame :645-670), one warp, in paint order
__device__ void paint_hud_indicators(const RasterSmem& S, const uint8_t* glyphs, const uint8_t* G, uint8_t* img, int lane) {
    const double W = CAR_W, H = CAR_H, s = W / 40.0, h = H / 40.0;
    hud_rect(img, 5 * s, H - h, s, h * (-0.02 * S.hud_vals[0]), G[G_BLUE], lane, 32);
    __syncwarp();
    for (int k = 0; k < 4; ++k) {
        hud_rect(img, (7 + k) * s, H - h, s, h * (-0.01 * S.hud_vals[1 + k]), k < 2 ? G[G_BLUE] : G[G_BLUE2], lane, 32);
        __syncwarp();
    }
    hud_rect(img, 20 * s, H - 2 * h, s * (10.0 * S.hud_vals[5]), 2 * h, G[G_GREEN], lane, 32);
    __syncwarp();
    hud_rect(img, 30 * s, H - 2 * h, s * (0.8 * S.hud_vals[6]), 2 * h, G[G_RED], lane, 32);
    __syncwarp();
    if (glyphs != nullptr) {     // draw_text("%05.0f" % reward) at (W/100, H - H/20): lane = one pixel of the 4x8 glyph cell
        // "%05.0f": round half to even, sign kept for negative values, zero padded to width 5
        const double rv = S.hud_vals[7];
        double mag = rint(fabs(rv));
        char digits[24];
        int nd = 0;
        if (mag == 0) digits[nd++] = 0;
        {   // integer digits (exact, no fp64 divisions); a reward is bounded by 1000 + 0.1 per step, far below 2^32
            unsigned int mi = (unsigned int)fmin(mag, 4.0e9);
            while (mi != 0u) { const unsigned int qi = mi / 10u; digits[nd++] = (char)(mi - qi * 10u); mi = qi; }
        }
        const int neg = (rv < 0 || (rv == 0 && signbit(rv))) ? 1 : 0;
        const int body = nd + neg;
        int pen = (int)(W / 100);
        const int y0 = (int)(H - H / 20);
        const int pad = 5 > body ? 5 - body : 0;
        const int gx = lane & 3, gy = lane >> 2;
        for (int i = 0; i < body + pad; ++i) {
            int gi;
            if (neg && i == 0) gi = 10;
            else if (i < neg + pad) gi = 0;
            else gi = digits[nd - 1 - (i - neg - pad)];
            if (glyphs[(gi * 8 + gy) * 4 + gx]) {
                const int px = pen + gx, py = y0 + gy;
                if (px >= 0 && px < CAR_W && py >= 0 && py < CAR_H) img[py * CAR_W + px] = G[G_TEXT];
            }
            pen += glyphs[11 * 8 * 4 + gi];
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(RASTER_THREADS, 5)
car_render_kernel(CarDev p, int only_done, int which, uint8_t* __restrict__ obs, uint8_t* __restrict__ term_obs) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    RasterSmem& S = *reinterpret_cast<RasterSmem*>(smem_raw);
    const int frame = blockIdx.x;                     // env * players + player
    const int e = (p.players == 2) ? frame >> 1 : frame, pi = (p.players == 2) ? frame & 1 : 0;   // players is 1 or 2
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (only_done && !p.env_done[e]) return;
    if (which != 0 && (p.deferred[e] != 0) != (which == 2)) return;
    const CarHullConst* K = p.consts;
    const uint8_t* G = K->gray;
    const int* checker = K->checker;
    const FrameMap fm = p.frame_map[frame];           // written by car_frame_setup_kernel
    const double obs_scale = car_obs_scale();
    const int slot = car_slot(p, e);
    const int n_track = p.n_track[slot];
    const CarTile* tiles = p.tiles + (size_t)slot * CAR_MAX_TRACK;
    const int C = p.c;
    // stack mode: the frames live in an internal ring [frame][C] and the observation (C channels, oldest first) is
    // rewritten every step.  Ring mode: the observation buffer itself is a double-write ring of 2C slots (the new frame
    // goes to slots k and k + C, the caller looks at slots k+1 .. k+C), so nothing is copied.
    const bool ringm = p.ring_mode != 0;
    uint8_t* ring = ringm ? nullptr : p.ring + (size_t)frame * C * CAR_PIX;
    const bool fill_all = only_done != 0 || (ringm ? p.fill_all != 0 : p.ring_pos[e] < 0);
    int newest = C - 1;
    if (ringm) newest = p.ring_phase;
    else if (!fill_all) { newest = p.ring_pos[e] + 1; if (newest >= C) newest = 0; }
    uint8_t* out = obs + (size_t)frame * (ringm ? 2 * C : C) * CAR_PIX;
    uint8_t* tout = (term_obs != nullptr && !only_done && p.env_done[e]) ? term_obs + (size_t)frame * C * CAR_PIX : nullptr;

    if (tid < p.players * 40) (&S.car_body[0][0])[tid] = p.body[(size_t)e * p.players * 40 + tid];   // [player][40], contiguous on both sides
    if (tid < (2 * CAR_W + 2 * CAR_H) / 4) reinterpret_cast<uint32_t*>(S.chk_x)[tid] = 0u;     // chk_x and chk_y are adjacent
    for (int i = tid; i < WALK_CELLS * MASK_WORDS; i += RASTER_THREADS) (&S.cell_mask[0][0])[i] = 0u;
    if (tid == 255) {
        S.n_poly = 0; S.pool_used = 0; S.overflow = 0; S.copy_next = 0; S.hud_late = 0;
        S.tiles = tiles; S.env_spans = p.tile_spans + (size_t)slot * CAR_MAX_TRACK * CAR_SPAN_ROWS;
    }
    if (tid >= 240 && tid < 248) {      // HUD inputs (render_indicators_for_pygame :645-670)
        const int k = tid - 240;
        const float* b = p.body + (size_t)frame * 40;
        const double* wd = p.wheel + (size_t)frame * 8;
        double v;
        if (k == 0) { const double vx = (double)b[3], vy = (double)b[4]; v = sqrt(vx * vx + vy * vy); }
        else if (k <= 4) v = wd[k - 1];
        else if (k == 5) v = (double)(b[8 + 2] - b[2]);     // wheels[0].joint.angle
        else if (k == 6) v = (double)b[5];                  // hull.angularVelocity
        else v = p.reward[2 * (size_t)frame];
        S.hud_vals[k] = v;
    }
    __syncthreads();
    // ---- checker squares (:733-746) are axis-aligned in the road map: mark the crop columns / rows that lie in one
    //      (20 squares per axis, each ~29 px wide; the tables were cleared above) ----
    if (tid >= 192 && tid < 192 + 40) {
        const int q = tid - 192, is_y = q >= 20;
        const int lo = checker[2 * q], hi = checker[2 * q + 1], base = is_y ? fm.ry : fm.rx;
        uint8_t* tab = is_y ? S.chk_y : S.chk_x;
        for (int v = max(lo - base, 0); v <= min(hi - base, 2 * CAR_W - 1); ++v) tab[v] = 1;
    }
    // ---- polygons of the frame, one thread each.  Road: paint order is tile n-1 .. 0, each followed by its kerb
    //      (:399-445).  Cars (Car.draw_for_pygame): for k in cars: wheels, then hull fixtures; b2Vec2 fp32 arithmetic:
    //      path = -scale * (tmp * ((trans * v) - offset)) + (W/2, H/2), truncated to int by pygame.  Higher key wins. ----
    {
        const int nc = p.frame_ncand[frame], per_car = 8;   // 4 wheels + 4 hull fixtures
        if (tid < nc) {
            const int t = p.frame_cand[(size_t)frame * CAR_MAX_CAND + tid];
            const CarTile T = tiles[t];
            const unsigned int order = 2u * (unsigned)(n_track - 1 - t) + 1u;
            int vx[8], vy[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { vx[i] = (i < 5) ? T.mx[i] : 0; vy[i] = (i < 5) ? T.my[i] : 0; }
            add_polygon(S, fm, vx, vy, 5, (order << 8) | G[G_ROAD0 + t % 3], false, 0, 0x40000000u | (unsigned)t);
            if (T.flags & 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { vx[i] = T.kmx[i]; vy[i] = T.kmy[i]; }
                add_polygon(S, fm, vx, vy, 4, ((order + 1u) << 8) | ((T.flags & 4) ? G[G_KERB_W] : G[G_KERB_R]), false, 0, 0x40010000u | (unsigned)t);
            }
        } else if (tid >= RASTER_THREADS - p.players * per_car) {
            const int q = RASTER_THREADS - 1 - tid;
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
                    const float sxp = (float)((double)rx2 * -obs_scale) + (float)(CAR_W / 2.0);
                    const float syp = (float)((double)ry2 * -obs_scale) + (float)(CAR_H / 2.0);
                    vx[i] = (int)sxp; vy[i] = (int)syp;
                }
            }
            const uint8_t g = (part < 4) ? G[G_WHEEL] : ((ck == pi) ? G[G_OWN] : G[G_OTHER]);
            add_polygon(S, fm, vx, vy, n, (order << 8) | g, true, q, 0u);
        }
    }
    // ---- meanwhile (nothing here touches what the polygon threads write): warp HUD_WARP paints the HUD bar -- rows 86..95,
    //      which the pixel pass below never writes -- and every warp, as soon as it is free, copies chunks of the C - 1
    //      older frames ring -> observation.  FrameStack: the new frame enters the ring; the observation is the ring
    //      oldest -> newest.  After a reset (only_done pass, or the very first render) every slot holds the reset frame.
    //      Output layout: [env][players * C][96][96]: player-major channel blocks (FlattenMultiAgentObservation
    //      concatenates the players' stacks on the channel axis), oldest frame first within a player. ----
    uint8_t* img = S.img;
    __syncwarp();
    if (warp == HUD_WARP) {
        // the indicators are painted after the scene in the reference (render_indicators_for_pygame :645-670); they stay inside
        // the bar unless a vertical one is taller than 7 px (speed >= 146, wheel omega >= 292): then they wait for the scene
        const double H = CAR_H, h = H / 40.0;
        bool late = false;
        for (int k = 0; k < 5; ++k) {
            const int Y = (int)(H - h), Hh = (int)(h * ((k == 0 ? -0.02 : -0.01) * S.hud_vals[k]));
            late = late || min(Y, Y + Hh - 1) < HUD_TOP;
        }
        const uint32_t g4 = 0x01010101u * (uint32_t)G[G_HUD];
        for (int q = lane; q < (CAR_H - HUD_TOP) * CAR_W / 16; q += 32) reinterpret_cast<uint4*>(img + HUD_TOP * CAR_W)[q] = make_uint4(g4, g4, g4, g4);
        __syncwarp();
        if (late) { if (lane == 0) S.hud_late = 1; }
        else paint_hud_indicators(S, p.glyphs, G, img, lane);
        __syncwarp();
    }
    if (!fill_all && (!ringm || tout != nullptr)) {                // ring mode: only the terminal observation is materialised
        constexpr int PER_SLOT = CAR_PIX / 16 / 64;                // chunks of 64 uint4 (two per lane) per frame: 9
        const int n_chunks = (C - 1) * PER_SLOT;
        for (;;) {
            int ch = 0;
            if (lane == 0) ch = atomicAdd(&S.copy_next, 1);
            ch = __shfl_sync(0xffffffffu, ch, 0);
            if (ch >= n_chunks) break;
            const int sl = ch / PER_SLOT, q = (ch % PER_SLOT) * 64 + lane;       // slot sl of the output = the sl-th oldest frame
            int rs = newest + 1 + sl;                             // < 2 C
            if (!ringm && rs >= C) rs -= C;
            const uint4* rsrc = reinterpret_cast<const uint4*>((ringm ? out : ring) + (size_t)rs * CAR_PIX);
            const uint4 v0 = rsrc[q], v1 = rsrc[q + 32];
            if (!ringm) {
                uint4* dst = reinterpret_cast<uint4*>(out + (size_t)sl * CAR_PIX);
                dst[q] = v0; dst[q + 32] = v1;
            }
            if (tout) { uint4* tdst = reinterpret_cast<uint4*>(tout + (size_t)sl * CAR_PIX); tdst[q] = v0; tdst[q + 32] = v1; }
        }
    }
    __syncthreads();
    if (tid == 0 && S.overflow != 0) atomicAdd(p.overrun + (S.overflow == 2 ? 1 : 2), 1);   // [1] polygons dropped, [2] frames with a full span pool (slow path)
    // ---- span tables: one (polygon, row) per thread and pass ----
    {
        const int total = min(S.pool_used, POOL_ROWS);
        for (int i = tid; i < total; i += RASTER_THREADS) {
            const int id = S.row_owner[i];
            const PolyMeta m = S.meta[id];
            if (m.off == NO_TABLE || i < (int)m.off || i >= (int)m.off + m.rows) continue;   // tail of a polygon that did not fit
            S.spans[i] = poly_row_spans(S, id, m, m.miny + (i - (int)m.off));   // road: copied from the per-reset tables
        }
    }
    __syncthreads();
    // ---- pixels above the HUD bar ----
    if (S.overflow == 0) walk_cells<false>(S, fm, G[G_GRASS], G[G_CHECK], warp, lane);
    else walk_cells<true>(S, fm, G[G_GRASS], G[G_CHECK], warp, lane);
    __syncthreads();
    if (S.hud_late) {                                              // uniform over the CTA
        if (warp == HUD_WARP) paint_hud_indicators(S, p.glyphs, G, img, lane);
        __syncthreads();
    }
    const uint4* src = reinterpret_cast<const uint4*>(img);
    if (fill_all && ringm) {
        for (int sl = 0; sl < 2 * C; ++sl) {
            uint4* dst = reinterpret_cast<uint4*>(out + (size_t)sl * CAR_PIX);
            for (int q = tid; q < CAR_PIX / 16; q += RASTER_THREADS) dst[q] = src[q];
        }
    } else if (fill_all) {
        for (int sl = 0; sl < C; ++sl) {
            uint4* rdst = reinterpret_cast<uint4*>(ring + (size_t)sl * CAR_PIX);
            uint4* dst = reinterpret_cast<uint4*>(out + (size_t)sl * CAR_PIX);
            for (int q = tid; q < CAR_PIX / 16; q += RASTER_THREADS) { const uint4 vv = src[q]; rdst[q] = vv; dst[q] = vv; }
        }
    } else {
        // stack: internal ring slot + newest channel.  ring: slots k and k + C of the observation ring
        uint4* rdst = reinterpret_cast<uint4*>((ringm ? out : ring) + (size_t)newest * CAR_PIX);
        uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(ringm ? newest + C : C - 1) * CAR_PIX);
        uint4* tdst = tout ? reinterpret_cast<uint4*>(tout + (size_t)(C - 1) * CAR_PIX) : nullptr;
        for (int q = tid; q < CAR_PIX / 16; q += RASTER_THREADS) {
            const uint4 vv = src[q];
            rdst[q] = vv;
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

size_t car_frame_map_bytes() { return sizeof(FrameMap); }

cudaError_t car_raster_init() {
    return cudaFuncSetAttribute(car_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RasterSmem));
}

cudaError_t launch_car_render(const CarDev& p, int only_done, int which, int advance, uint8_t* obs, uint8_t* term_obs, cudaStream_t s) {
    car_frame_setup_kernel<<<(p.n * p.players + RASTER_WARPS - 1) / RASTER_WARPS, RASTER_THREADS, 0, s>>>(p, only_done, which);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    car_render_kernel<<<p.n * p.players, RASTER_THREADS, sizeof(RasterSmem), s>>>(p, only_done, which, obs, term_obs);
    e = cudaGetLastError();
    if (e != cudaSuccess || !advance || p.ring_mode) return e;   // ring mode: the phase is advanced on the host, per step
    car_ring_advance_kernel<<<(p.n + 127) / 128, 128, 0, s>>>(p, only_done);
    return cudaGetLastError();
}

}  // namespace crl
