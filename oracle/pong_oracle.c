/*
 * pong_oracle.c -- CPU restatement of the reference's vectorised Pong stepping path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle and the `cpu_baseline`
 * of bench.py.  Nothing in the product package may link, import or execute it;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs do.
 *
 * Parity pin: the reference ships no golden vectors, so the pin
 * is the reference's own Python files executed in the build container under
 * stand-in gym/pygame modules (oracle/ref_loader.py) with the real cv2 4.13;
 * their outputs are committed under tests/golden/ (generator: oracle/gen_golden.py)
 * and tests/test_oracle_golden.py checks this file against them bit for bit.
 * Third-party arithmetic restated here (pygame Rect truncation, cv2 gray/area) is
 * listed per function.  Scoreboard glyph pixels are DATA (the atlas) -- see
 * DESIGN.md "scoreboard atlas".
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/competitive_rl/).
 *
 * Build: gcc -O2 -ffp-contract=off -pthread -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: cv2's area resize is un-fused mul-then-add in fp32.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SCREEN_W 160
#define SCREEN_H 210
#define FRAME_BYTES (SCREEN_W * SCREEN_H * 3)
#define ATLAS_ROWS 34 /* rows above the arena: 0..33 */
#define ATLAS_SCORES 22
#define CHEAT_CODES 999 /* pong/base_pong_env.py:9 */

/* PongGame.__init__ geometry, pong/base_pong_env.py:158-211 with the env ctor
 * arguments of :27-36 (window 160x210, ball_speed=4, bat_speed=4). */
#define ARENA_LEFT 0
#define ARENA_RIGHT 160
#define ARENA_TOP 34
#define ARENA_BOTTOM 194 /* Rect(0, 34, 160, 160): height = window WIDTH, :275-276 */
#define ARENA_CENTERY 114
#define BALL_SIZE 4
#define BALL_X0 78
#define BALL_Y0 112
#define BAT_W 5
#define BAT_H 15
#define BAT_Y0 107
#define LEFT_BAT_X 16
#define RIGHT_BAT_X 139
#define SPEED 4
#define MAX_STEP_PER_ROUND 10000

typedef struct {
    int ball_x, ball_y;
    double vx, vy;
    int left_y, right_y;
    int left_move, right_move; /* Bat._current_move */
    int score_left, score_right;
    int num_rounds, num_steps;
} Game;

typedef struct {
    int dst, src;
    float alpha;
} Tap;

typedef struct {
    Tap* taps;
    int n;
} Tab;

typedef struct {
    Game g;
    /* MaxAndSkipEnv._obs_buffer: [agent][slot][210*160*3], utils/atari_wrappers.py:106-115 */
    uint8_t* obs_buffer;
    /* FrameStack deque, oldest first: [agent][n_frames][dim*dim], :222-259 */
    uint8_t* frames;
    int clip_steps; /* ClipRewardEnv._steps, :166-181 */
    int serve_count;
    uint64_t rng;
} Env;

typedef struct {
    int n, n_agents, dim, n_stack, c; /* c = max(n_stack,1) channels */
    int double_player, max_rounds, render;
    const uint8_t* atlas; /* [22][22][34][160][3] */
    const double* serves; /* [n][K][2] or NULL */
    int K;
    Tab xtab, ytab;
    Env* envs;
    int serve_overrun;
} Vec;

/* ------------------------------------------------------------------------- */
/* serve randomness                                                           */

static double rng_uniform01(uint64_t* s) {
    /* splitmix64; only used when no serve table is injected (CPU baseline timing) */
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

/* Ball.reset, pong/base_pong_env.py:314-320.  Draw order uniform, choice, choice;
 * in validation mode the three draws are replaced by the next table entry
 * serves[env][k] = (vx, vy) (oracle/ref_loader.py::ServeInjector does the same to
 * the reference). */
static void ball_reset(Vec* v, Env* e, int env_idx) {
    Game* g = &e->g;
    g->ball_x = BALL_X0;
    g->ball_y = BALL_Y0;
    if (v->serves) {
        int k = e->serve_count;
        if (k >= v->K) {
            v->serve_overrun = 1;
            k = v->K - 1;
        }
        g->vx = v->serves[((size_t)env_idx * v->K + k) * 2 + 0];
        g->vy = v->serves[((size_t)env_idx * v->K + k) * 2 + 1];
    } else {
        double lo = SPEED * 0.3, hi = SPEED;
        double vy0 = lo + (hi - lo) * rng_uniform01(&e->rng);
        g->vx = rng_uniform01(&e->rng) < 0.5 ? -(double)SPEED : (double)SPEED;
        g->vy = rng_uniform01(&e->rng) < 0.5 ? -vy0 : vy0;
    }
    e->serve_count++;
}

/* ------------------------------------------------------------------------- */
/* game core                                                                  */

/* auto_action, pong/base_pong_env.py:457-471 */
static int auto_action(double ball_speed_x, int rect_center_y, int ball_center_y, int arena_center_y) {
    int direction = 0;
    if (ball_speed_x < 0) {
        if (rect_center_y < arena_center_y) direction = 1;
        else if (rect_center_y > arena_center_y) direction = -1;
    } else if (ball_speed_x > 0) {
        if (rect_center_y < ball_center_y) direction = 1;
        else direction = -1;
    }
    return direction;
}

/* Bat.move, pong/base_pong_env.py:412-418 (AutoBat.move :445-454 differs only in
 * where the direction comes from). */
static void bat_move(int* y, int* current_move, int direction) {
    *current_move = direction * SPEED;
    *y += *current_move;
    if (*y + BAT_H > ARENA_BOTTOM) *y += ARENA_BOTTOM - (*y + BAT_H);
    else if (*y < ARENA_TOP) *y += ARENA_TOP - *y;
}

/* pygame 1.9.x Rect attribute store: C (int) cast == truncation toward zero
 * [third-party behaviour, restated]. */
static int rect_store(double v) { return (int)v; }

/* Ball.move + _bounce, pong/base_pong_env.py:325-361, 375-381 */
static void ball_move(Game* g) {
    int prev_left = g->ball_x, prev_right = g->ball_x + BALL_SIZE;
    int rb_left = RIGHT_BAT_X, lb_right = LEFT_BAT_X + BAT_W;
    double y_on_right_bat = (double)(rb_left - prev_right) / g->vx * g->vy + (double)g->ball_y;
    double y_on_left_bat = (double)(lb_right - prev_left) / g->vx * g->vy + (double)g->ball_y;
    g->ball_x = rect_store((double)g->ball_x + g->vx);
    g->ball_y = rect_store((double)g->ball_y + g->vy);
    if (g->vy < 0 && g->ball_y <= ARENA_TOP) {
        g->vy *= -1;
        g->vx += 0;
        g->ball_y = ARENA_TOP;
    } else if (g->vy > 0 && g->ball_y + BALL_SIZE >= ARENA_BOTTOM) {
        g->vy *= -1;
        g->vx += 0;
        g->ball_y = ARENA_BOTTOM - BALL_SIZE;
    } else if (g->vx < 0 && g->ball_x <= lb_right && y_on_left_bat + BALL_SIZE >= g->left_y &&
               y_on_left_bat <= g->left_y + BAT_H && prev_left > lb_right) {
        g->vx *= -1;
        g->vy += g->left_move * 0.7;
        g->ball_x = lb_right;
        g->ball_y = rect_store(y_on_left_bat);
    } else if (g->vx > 0 && g->ball_x + BALL_SIZE >= rb_left && y_on_right_bat + BALL_SIZE >= g->right_y &&
               y_on_right_bat <= g->right_y + BAT_H && prev_right < rb_left) {
        g->vx *= -1;
        g->vy += g->right_move * 0.7;
        g->ball_x = rb_left - BALL_SIZE;
        g->ball_y = rect_store(y_on_right_bat);
    }
}

/* PongGame._reset_round, pong/base_pong_env.py:247-250 */
static void reset_round(Vec* v, Env* e, int idx) {
    ball_reset(v, e, idx);
    e->g.num_rounds += 1;
    e->g.num_steps = 0;
}

static void bats_reset(Game* g) { /* Bat.reset :420-422 (current_move is NOT cleared) */
    g->left_y = BAT_Y0;
    g->right_y = BAT_Y0;
}

/* PongGame.reset_game, pong/base_pong_env.py:252-257 */
static void reset_game(Vec* v, Env* e, int idx) {
    e->g.score_left = e->g.score_right = 0;
    reset_round(v, e, idx);
    bats_reset(&e->g);
    e->g.num_rounds = 0;
}

/* PongGame.step, pong/base_pong_env.py:213-245.  right_dir is ignored for the
 * single-player game (AutoBat). */
static int game_step(Vec* v, Env* e, int idx, int left_dir, int right_dir, int rewards[2]) {
    Game* g = &e->g;
    g->num_steps += 1;
    bat_move(&g->left_y, &g->left_move, left_dir);
    if (v->double_player) {
        bat_move(&g->right_y, &g->right_move, right_dir);
    } else {
        int d = auto_action(g->vx, g->right_y + (BAT_H >> 1), g->ball_y + (BALL_SIZE >> 1), ARENA_CENTERY);
        bat_move(&g->right_y, &g->right_move, d);
    }
    ball_move(g);
    rewards[0] = rewards[1] = 0;
    if (g->ball_x < ARENA_LEFT) {
        g->score_right += 1;
        rewards[0] = -1;
        rewards[1] = 1;
        reset_round(v, e, idx);
        bats_reset(g);
    } else if (g->ball_x + BALL_SIZE > ARENA_RIGHT) {
        g->score_left += 1;
        rewards[0] = 1;
        rewards[1] = -1;
        reset_round(v, e, idx);
        bats_reset(g);
    } else if (g->num_steps > MAX_STEP_PER_ROUND) {
        reset_round(v, e, idx);
        bats_reset(g);
    }
    return g->num_rounds >= v->max_rounds;
}

/* PongDoublePlayerEnv._step action decode, pong/base_pong_env.py:113-142;
 * PongSinglePlayerEnv._step :41-46. */
static int env_step_raw(Vec* v, Env* e, int idx, const int32_t* action, int rewards[2]) {
    static const int BAT_DIRECTIONS[3] = {-1, 0, 1};
    Game* g = &e->g;
    int left_dir, right_dir = 0;
    if (v->double_player) {
        int la = action[0], ra = action[1];
        if (ra == CHEAT_CODES)
            right_dir = auto_action(g->vx, g->right_y + (BAT_H >> 1), g->ball_y + (BALL_SIZE >> 1), ARENA_CENTERY);
        else
            right_dir = BAT_DIRECTIONS[ra];
        if (la == CHEAT_CODES)
            left_dir = auto_action(-g->vx, g->left_y + (BAT_H >> 1), g->ball_y + (BALL_SIZE >> 1), ARENA_CENTERY);
        else
            left_dir = BAT_DIRECTIONS[la];
    } else {
        left_dir = BAT_DIRECTIONS[action[0]];
    }
    return game_step(v, e, idx, left_dir, right_dir, rewards);
}

/* ------------------------------------------------------------------------- */
/* renderer                                                                   */

static void fill_rect(uint8_t* rgb, int x, int y, int w, int h, uint8_t c) {
    /* pygame.draw.rect: filled, clipped to the surface [third-party, restated] */
    int x0 = x < 0 ? 0 : x, y0 = y < 0 ? 0 : y;
    int x1 = x + w > SCREEN_W ? SCREEN_W : x + w, y1 = y + h > SCREEN_H ? SCREEN_H : y + h;
    for (int r = y0; r < y1; ++r)
        if (x1 > x0) memset(rgb + ((size_t)r * SCREEN_W + x0) * 3, c, (size_t)(x1 - x0) * 3);
}

/* _get_screen_img, pong/base_pong_env.py:66-74: Arena.draw (:278-280) fills WHITE
 * then the arena rect BLACK; Ball.draw (:322-323), Bat.draw (:406-410) WHITE;
 * Scoreboard.draw (:480-487) blits "Score = %d : %d" at (20, 8).  The blitted
 * text only touches rows 8..33, where everything underneath is white, so the
 * composite equals the atlas strip for (score_left, score_right). */
static void render_frame(const Vec* v, const Game* g, uint8_t* rgb) {
    memset(rgb, 255, FRAME_BYTES);
    fill_rect(rgb, 0, ARENA_TOP, SCREEN_W, SCREEN_W, 0);
    fill_rect(rgb, g->ball_x, g->ball_y, BALL_SIZE, BALL_SIZE, 255);
    fill_rect(rgb, LEFT_BAT_X, g->left_y, BAT_W, BAT_H, 255);
    fill_rect(rgb, RIGHT_BAT_X, g->right_y, BAT_W, BAT_H, 255);
    if (v->atlas) {
        int l = g->score_left, r = g->score_right;
        if (l < ATLAS_SCORES && r < ATLAS_SCORES)
            memcpy(rgb, v->atlas + ((size_t)l * ATLAS_SCORES + r) * ATLAS_ROWS * SCREEN_W * 3,
                   (size_t)ATLAS_ROWS * SCREEN_W * 3);
    }
}

/* _get_screen_img_double_player, pong/base_pong_env.py:149-155:
 * new_flip[25:] = new_flip[25:, ::-1] -- rows 0..24 are NOT mirrored. */
static void flip_for_agent1(const uint8_t* rgb, uint8_t* out) {
    memcpy(out, rgb, (size_t)25 * SCREEN_W * 3);
    for (int r = 25; r < SCREEN_H; ++r)
        for (int c = 0; c < SCREEN_W; ++c)
            memcpy(out + ((size_t)r * SCREEN_W + c) * 3, rgb + ((size_t)r * SCREEN_W + (SCREEN_W - 1 - c)) * 3, 3);
}

/* ------------------------------------------------------------------------- */
/* cv2 arithmetic [third-party, restated; pinned against cv2 4.13 in
 * tests/test_oracle_cv2.py]                                                   */

/* cv2.cvtColor(COLOR_RGB2GRAY) on uint8: 15-bit fixed point, SURVEY.md A.1 */
static void cvt_gray(const uint8_t* rgb, uint8_t* gray, int npix) {
    for (int i = 0; i < npix; ++i)
        gray[i] = (uint8_t)((rgb[3 * i] * 9798 + rgb[3 * i + 1] * 19235 + rgb[3 * i + 2] * 3735 + 16384) >> 15);
}

/* OpenCV computeResizeAreaTab (imgproc/resize.cpp), SURVEY.md A.2 */
static Tab compute_area_tab(int ssize, int dsize) {
    Tab t;
    t.taps = (Tap*)malloc(sizeof(Tap) * (size_t)(ssize + 2 * dsize + 4));
    t.n = 0;
    double scale = (double)ssize / dsize;
    for (int dx = 0; dx < dsize; ++dx) {
        double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        double cell = scale < ssize - fsx1 ? scale : ssize - fsx1;
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        if (sx2 > ssize - 1) sx2 = ssize - 1;
        if (sx1 > sx2) sx1 = sx2;
        if (sx1 - fsx1 > 1e-3) t.taps[t.n++] = (Tap){dx, sx1 - 1, (float)((sx1 - fsx1) / cell)};
        for (int sx = sx1; sx < sx2; ++sx) t.taps[t.n++] = (Tap){dx, sx, (float)(1.0 / cell)};
        if (fsx2 - sx2 > 1e-3) {
            double w = fsx2 - sx2;
            if (w > 1.0) w = 1.0;
            if (w > cell) w = cell;
            t.taps[t.n++] = (Tap){dx, sx2, (float)(w / cell)};
        }
    }
    return t;
}

/* cv2.resize(uint8 1-channel, INTER_AREA) with a non-integer scale:
 * ResizeArea_<uchar,float>: per source row a horizontal pass into buf (mul, add),
 * rows accumulated into sum with beta (mul, add), saturate_cast<uchar>(rint). */
static void resize_area(const Vec* v, const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
    float* buf = (float*)malloc(sizeof(float) * (size_t)dw);
    float* sum = (float*)malloc(sizeof(float) * (size_t)dw);
    const Tab* xt = &v->xtab;
    const Tab* yt = &v->ytab;
    (void)sh;
    int prev_dy = yt->taps[0].dst;
    for (int dx = 0; dx < dw; ++dx) sum[dx] = 0.f;
    for (int j = 0; j < yt->n; ++j) {
        int dy = yt->taps[j].dst, sy = yt->taps[j].src;
        float beta = yt->taps[j].alpha;
        for (int dx = 0; dx < dw; ++dx) buf[dx] = 0.f;
        const uint8_t* S = src + (size_t)sy * sw;
        for (int k = 0; k < xt->n; ++k) {
            int dxn = xt->taps[k].dst;
            float p = (float)S[xt->taps[k].src] * xt->taps[k].alpha;
            buf[dxn] = buf[dxn] + p;
        }
        if (dy != prev_dy) {
            for (int dx = 0; dx < dw; ++dx) {
                long r = lrintf(sum[dx]); /* round-half-even (default FE_TONEAREST) */
                dst[(size_t)prev_dy * dw + dx] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
                sum[dx] = beta * buf[dx];
            }
            prev_dy = dy;
        } else {
            for (int dx = 0; dx < dw; ++dx) {
                float p = beta * buf[dx];
                sum[dx] = sum[dx] + p;
            }
        }
    }
    for (int dx = 0; dx < dw; ++dx) {
        long r = lrintf(sum[dx]);
        dst[(size_t)prev_dy * dw + dx] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
    }
    free(buf);
    free(sum);
}

/* WarpFrame.parse_single_frame, utils/atari_wrappers.py:215-219 */
static void warp_frame(const Vec* v, const uint8_t* rgb, uint8_t* out) {
    uint8_t* gray = (uint8_t*)malloc((size_t)SCREEN_W * SCREEN_H);
    cvt_gray(rgb, gray, SCREEN_W * SCREEN_H);
    resize_area(v, gray, SCREEN_W, SCREEN_H, out, v->dim, v->dim);
    free(gray);
}

/* exported for the cv2 pin test: gray + area resize of one RGB frame */
void pong_oracle_warp(const uint8_t* rgb, int dim, uint8_t* out) {
    Vec v;
    memset(&v, 0, sizeof v);
    v.dim = dim;
    v.xtab = compute_area_tab(SCREEN_W, dim);
    v.ytab = compute_area_tab(SCREEN_H, dim);
    warp_frame(&v, rgb, out);
    free(v.xtab.taps);
    free(v.ytab.taps);
}

/* generic single-channel area resize for the cv2 pin test */
void pong_oracle_resize_area(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
    Vec v;
    memset(&v, 0, sizeof v);
    v.xtab = compute_area_tab(sw, dw);
    v.ytab = compute_area_tab(sh, dh);
    resize_area(&v, src, sw, sh, dst, dw, dh);
    free(v.xtab.taps);
    free(v.ytab.taps);
}

/* ------------------------------------------------------------------------- */
/* wrapper stack                                                              */

static size_t frame_px(const Vec* v) { return (size_t)v->dim * v->dim; }

static uint8_t* buf_slot(const Vec* v, Env* e, int agent, int slot) {
    (void)v;
    return e->obs_buffer + ((size_t)agent * 2 + slot) * FRAME_BYTES;
}

/* raw observation of the env: 1 frame (single) or (frame, flipped frame) */
static void raw_obs(const Vec* v, Env* e, uint8_t* out0, uint8_t* out1) {
    render_frame(v, &e->g, out0);
    if (v->double_player) flip_for_agent1(out0, out1);
}

/* FrameStack: append to the deque (maxlen n), utils/atari_wrappers.py:252-255 */
static void stack_push(const Vec* v, Env* e, int agent, const uint8_t* frame) {
    size_t px = frame_px(v);
    uint8_t* base = e->frames + (size_t)agent * v->c * px;
    if (v->c > 1) memmove(base, base + px, (size_t)(v->c - 1) * px);
    memcpy(base + (size_t)(v->c - 1) * px, frame, px);
}

/* env.reset() through the whole stack: FrameStack.reset (:246-250, n copies) ->
 * ClipRewardEnv.reset (:171-173) -> WarpFrame.observation of the RAW frame
 * (no max-pool; MaxAndSkipEnv.reset :162-163 passes through and does NOT clear
 * _obs_buffer) -> Pong*Env._reset -> reset_game. */
static void env_reset(Vec* v, Env* e, int idx, uint8_t* scratch /* 2*FRAME_BYTES + dim*dim */) {
    reset_game(v, e, idx);
    e->clip_steps = 0;
    if (!v->render) return;
    uint8_t* f0 = scratch;
    uint8_t* f1 = scratch + FRAME_BYTES;
    uint8_t* small = scratch + 2 * FRAME_BYTES;
    raw_obs(v, e, f0, f1);
    for (int a = 0; a < v->n_agents; ++a) {
        warp_frame(v, a == 0 ? f0 : f1, small);
        for (int k = 0; k < v->c; ++k) stack_push(v, e, a, small);
    }
}

/* copy the current stacked observation of env e into out[agent][idx][c][dim][dim]
 * (WrapPyTorch: HWC -> CHW, :35-37; channel 0 = oldest frame, :257-259) */
static void emit_obs(const Vec* v, Env* e, int idx, uint8_t* out) {
    size_t per = (size_t)v->c * frame_px(v);
    for (int a = 0; a < v->n_agents; ++a)
        memcpy(out + ((size_t)a * v->n + idx) * per, e->frames + (size_t)a * per, per);
}

/* MaxAndSkipEnv.step (:118-160) + WarpFrame + ClipRewardEnv.step (:175-181) +
 * FrameStack.step.  Returns done. */
static int env_step(Vec* v, Env* e, int idx, const int32_t* action, float* real_reward, uint8_t* scratch) {
    double total[2] = {0.0, 0.0};
    int done = 0;
    uint8_t* f0 = scratch;
    uint8_t* f1 = scratch + FRAME_BYTES;
    uint8_t* small = scratch + 2 * FRAME_BYTES;
    for (int i = 0; i < 4; ++i) {
        int rewards[2];
        done = env_step_raw(v, e, idx, action, rewards);
        if (v->render) {
            raw_obs(v, e, f0, f1); /* the reference renders on every sub-step */
            if (i == 2 || i == 3)
                for (int a = 0; a < v->n_agents; ++a)
                    memcpy(buf_slot(v, e, a, i - 2), a == 0 ? f0 : f1, FRAME_BYTES);
        }
        total[0] += rewards[0];
        total[1] += rewards[1];
        if (done) break;
    }
    if (v->render) {
        for (int a = 0; a < v->n_agents; ++a) {
            const uint8_t* b0 = buf_slot(v, e, a, 0);
            const uint8_t* b1 = buf_slot(v, e, a, 1);
            for (int k = 0; k < FRAME_BYTES; ++k) f0[k] = b0[k] > b1[k] ? b0[k] : b1[k];
            warp_frame(v, f0, small);
            stack_push(v, e, a, small);
        }
    }
    e->clip_steps += 1;
    real_reward[0] = (float)total[0];
    real_reward[1] = (float)total[1];
    return done;
}

/* ------------------------------------------------------------------------- */
/* vec env (C ABI for ctypes)                                                 */

void pong_oracle_destroy(Vec* v) {
    if (!v) return;
    for (int i = 0; i < v->n; ++i) {
        free(v->envs[i].obs_buffer);
        free(v->envs[i].frames);
    }
    free(v->envs);
    free(v->xtab.taps);
    free(v->ytab.taps);
    free(v);
}

/* Construction = make_env_a2c_atari thunk (utils/atari_wrappers.py:40-53): the
 * PongGame constructor consumes TWO serves (Ball.__init__ -> reset :312, then
 * reset_game :211). */
Vec* pong_oracle_create(int n_envs, int double_player, int resized_dim, int frame_stack, int max_rounds, int render,
                        const uint8_t* atlas, const double* serves, int K, uint64_t seed) {
    Vec* v = (Vec*)calloc(1, sizeof(Vec));
    v->n = n_envs;
    v->double_player = double_player;
    v->n_agents = double_player ? 2 : 1;
    v->dim = resized_dim;
    v->n_stack = frame_stack;
    v->c = frame_stack > 0 ? frame_stack : 1;
    v->max_rounds = max_rounds;
    v->render = render;
    v->atlas = atlas;
    v->serves = serves;
    v->K = K;
    v->xtab = compute_area_tab(SCREEN_W, resized_dim);
    v->ytab = compute_area_tab(SCREEN_H, resized_dim);
    v->envs = (Env*)calloc((size_t)n_envs, sizeof(Env));
    for (int i = 0; i < n_envs; ++i) {
        Env* e = &v->envs[i];
        e->rng = seed * 0x9E3779B97F4A7C15ull + (uint64_t)i * 0xD1B54A32D192ED03ull + 1;
        if (render) {
            e->obs_buffer = (uint8_t*)calloc((size_t)v->n_agents * 2, FRAME_BYTES); /* np.zeros */
            e->frames = (uint8_t*)calloc((size_t)v->n_agents * v->c, frame_px(v));
        }
        ball_reset(v, e, i);     /* Ball.__init__ */
        reset_game(v, e, i);     /* PongGame.__init__ tail */
    }
    return v;
}

int pong_oracle_serve_overrun(const Vec* v) { return v->serve_overrun; }

/* ---- tiny pthread parallel-for (the image has no libgomp) ---- */
typedef struct {
    Vec* v;
    int lo, hi;
    const int32_t* actions;
    uint8_t* obs_out;
    float* rew;
    uint8_t* done_out;
    int32_t* num_steps;
    float* real_reward;
    uint8_t* term_obs;
    int is_reset;
} Job;

static int g_threads = 1;
void pong_oracle_set_threads(int t) { g_threads = t < 1 ? 1 : t; }

static void* run_job(void* p) {
    Job* j = (Job*)p;
    Vec* v = j->v;
    int aw = v->double_player ? 2 : 1;
    uint8_t* scratch = (uint8_t*)malloc(2 * FRAME_BYTES + frame_px(v));
    for (int i = j->lo; i < j->hi; ++i) {
        Env* e = &v->envs[i];
        if (j->is_reset) {
            env_reset(v, e, i, scratch);
            if (v->render && j->obs_out) emit_obs(v, e, i, j->obs_out);
            continue;
        }
        float rr[2];
        int done = env_step(v, e, i, j->actions + (size_t)i * aw, rr, scratch);
        for (int a = 0; a < 2; ++a) {
            j->real_reward[i * 2 + a] = rr[a];
            j->rew[i * 2 + a] = (float)((rr[a] > 0) - (rr[a] < 0)); /* np.sign */
        }
        j->done_out[i] = (uint8_t)done;
        j->num_steps[i] = e->clip_steps;
        if (done) {
            if (v->render && j->term_obs) emit_obs(v, e, i, j->term_obs);
            env_reset(v, e, i, scratch);
        }
        if (v->render && j->obs_out) emit_obs(v, e, i, j->obs_out);
    }
    free(scratch);
    return NULL;
}

static void dispatch(Job proto) {
    int n = proto.v->n, T = g_threads > n ? (n > 0 ? n : 1) : g_threads;
    if (T <= 1) {
        proto.lo = 0;
        proto.hi = n;
        run_job(&proto);
        return;
    }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)T);
    Job* jobs = (Job*)malloc(sizeof(Job) * (size_t)T);
    for (int t = 0; t < T; ++t) {
        jobs[t] = proto;
        jobs[t].lo = (int)((long long)n * t / T);
        jobs[t].hi = (int)((long long)n * (t + 1) / T);
        pthread_create(&th[t], NULL, run_job, &jobs[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

/* DummyVecEnv.reset, utils/dummy_vec_env.py:71-75 */
void pong_oracle_reset(Vec* v, uint8_t* obs_out) {
    Job j;
    memset(&j, 0, sizeof j);
    j.v = v;
    j.obs_out = obs_out;
    j.is_reset = 1;
    dispatch(j);
}

/* DummyVecEnv.step_wait (utils/dummy_vec_env.py:51-63) / SubprocVecEnv _worker
 * (utils/subproc_vec_env.py:17-23): step, on done stash terminal_observation and
 * reset.  actions: [n][2] (double) or [n] (single).  rew: [n][2] clipped sign
 * rewards; real_reward: [n][2]; num_steps: [n] (info["num_steps"], counted
 * BEFORE the auto-reset zeroes it); term_obs (may be NULL): same layout as
 * obs_out, written only for envs with done=1.  Envs are independent, so the
 * loop is split over threads the way SubprocVecEnv splits it over processes. */
void pong_oracle_step(Vec* v, const int32_t* actions, uint8_t* obs_out, float* rew, uint8_t* done_out,
                      int32_t* num_steps, float* real_reward, uint8_t* term_obs) {
    Job j;
    memset(&j, 0, sizeof j);
    j.v = v;
    j.actions = actions;
    j.obs_out = obs_out;
    j.rew = rew;
    j.done_out = done_out;
    j.num_steps = num_steps;
    j.real_reward = real_reward;
    j.term_obs = term_obs;
    dispatch(j);
}

/* state[n][10] = ball_x, ball_y, vx, vy, left_y, right_y, score_l, score_r, rounds, steps */
void pong_oracle_get_state(const Vec* v, double* state) {
    for (int i = 0; i < v->n; ++i) {
        const Game* g = &v->envs[i].g;
        double* s = state + (size_t)i * 10;
        s[0] = g->ball_x; s[1] = g->ball_y; s[2] = g->vx; s[3] = g->vy;
        s[4] = g->left_y; s[5] = g->right_y; s[6] = g->score_left; s[7] = g->score_right;
        s[8] = g->num_rounds; s[9] = g->num_steps;
    }
}

/* raw 210x160x3 frame(s) of env i as the env would render them now (agent 0, agent 1) */
void pong_oracle_render_raw(Vec* v, int i, uint8_t* out0, uint8_t* out1) {
    render_frame(v, &v->envs[i].g, out0);
    if (out1) flip_for_agent1(out0, out1);
}

/* test hook: direct access to env i's game state (tests/test_oracle_golden.py) */
Game* pong_oracle_game_ptr(Vec* v, int i) { return &v->envs[i].g; }
