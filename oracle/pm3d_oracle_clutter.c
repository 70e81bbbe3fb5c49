/*
 * pm3d_oracle_clutter.c — CPU restatement of ClutterHandler.Update (game/clutter.go:25-117) and of the
 * Go standard library generator it draws from (math/rand: rand.New(rand.NewSource(..)), Rand.Seed,
 * Rand.Float64 — game/model/rand.go:9-15, game/clutter.go:75,83,93-94).
 *
 * TEST INFRASTRUCTURE ONLY (see pm3d_oracle.h).
 *   [V] follows a file under /root/reference (citation given)
 *   [R] recalled from a source that is not on disk:
 *       - Go's math/rand (rng.go: the additive lagged Fibonacci source, rngLen 607, rngTap 273, seeding through
 *         the Lehmer generator 48271 mod 2^31-1 and the rngCooked table; rand.go: Int63, Int31, Int31n, Intn,
 *         Float64 = float64(Int63()) / 2^63 with a resample on 1).  The 607-word rngCooked table is not
 *         recalled: it is REGENERATED the way the library's gen_cooked.go does (seed 1, 7.8e12 steps of the
 *         un-cooked generator), with a polynomial jump-ahead instead of 7.8e12 additions.
 *         PINNED by known answers every Go programmer has seen (tests/test_clutter_cpu.py): with seed 1,
 *         Intn(100) = 81 87 47 59 81 18 25 40 56 0, Int63() = 5577006791947779410 8674665223082153551,
 *         Float64() = 0.6046602879796196 0.9405090880450124 0.6645600532184904 0.4377141871869802
 *         0.4246374970712657, rngCooked[0] = -4181792142133755926, rngCooked[1] = -4576982950128230565.
 *       - geom.LineFromAngle(x, y, angle, length) = (x, y, x + length*cos(angle), y + length*sin(angle)) and
 *         geom.Radians(deg) = deg * math.Pi / 180 of github.com/harbdog/raycaster-go v1.12.0 (go.mod:13).
 *   [D] decisions of this repo:
 *       - cos/sin of the <= 21 sample angles are evaluated by the caller (Go's math.Sin/Cos are software
 *         routines that can differ from the C library's in the last ulp; the product takes them from the
 *         host the same way, like the dir/plane vectors of a view);
 *       - the handler's cache (spritesByPosition) only avoids re-seeding cells that stay in view: because the
 *         seed is Map.Seed + posId the sprites of a cell are the same whenever it is (re)loaded, so the state
 *         after Update is a pure function of the camera.  This restatement returns that state, in creation
 *         order (angle outer, distance inner, clutter index); the Go maps have no order.  posId =
 *         (x-1)*mapWidth + y is not unique on maps higher than wide: like the reference, a cell whose posId is
 *         already loaded is skipped — "already" meaning earlier in THIS walk.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pm3d_oracle.h"

/* ------------------------------------------------------------------ math/rand [R] */

#define RNG_LEN 607
#define RNG_TAP 273
#define INT32_MAX_ 2147483647

/* seedrand: x[n+1] = 48271 * x[n] mod (2^31 - 1), Schrage's method as in rng.go */
static int32_t seedrand(int32_t x) {
    const int32_t A = 48271, Q = 44488, R = 3399;
    int32_t hi = x / Q, lo = x % Q;
    x = A * lo - R * hi;
    if (x < 0) x += INT32_MAX_;
    return x;
}

/* c = a * b mod (x^607 - x^334 - 1), coefficients mod 2^64 */
static void poly_mul(const uint64_t *a, const uint64_t *b, uint64_t *out) {
    static _Thread_local uint64_t c[2 * RNG_LEN];
    memset(c, 0, sizeof c);
    for (int i = 0; i < RNG_LEN; i++) {
        if (!a[i]) continue;
        for (int j = 0; j < RNG_LEN; j++) c[i + j] += a[i] * b[j];
    }
    for (int d = 2 * RNG_LEN - 2; d >= RNG_LEN; d--) {
        c[d - RNG_LEN + (RNG_LEN - RNG_TAP)] += c[d];
        c[d - RNG_LEN] += c[d];
    }
    memcpy(out, c, RNG_LEN * sizeof(uint64_t));
}

/* the un-cooked generator of gen_cooked.go stepped n times from srand(1); direct form */
static void gen_direct(uint64_t *vec, uint64_t n) {
    int tap = 0, feed = RNG_LEN - RNG_TAP;
    for (uint64_t i = 0; i < n; i++) {
        if (--tap < 0) tap += RNG_LEN;
        if (--feed < 0) feed += RNG_LEN;
        vec[feed] = vec[feed] + vec[tap];
    }
}

/* the same n steps by jump-ahead: call k writes position (334 - k) mod 607 with y_k = y_{k-607} + y_{k-273},
 * so y_{m-606} = <x^m mod f, (y_{-606} .. y_0)> */
static void gen_jump(uint64_t *vec, uint64_t n) {
    uint64_t s[RNG_LEN], q[RNG_LEN], b[RNG_LEN], y[RNG_LEN];
    for (int j = 0; j < RNG_LEN; j++) {
        const int n0 = j - 606;
        s[j] = vec[n0 >= -272 ? 334 - n0 : -273 - n0];
    }
    memset(q, 0, sizeof q);
    memset(b, 0, sizeof b);
    q[0] = 1;
    b[1] = 1;
    for (uint64_t e = n; e; e >>= 1) {
        if (e & 1) poly_mul(q, b, q);
        poly_mul(b, b, b);
    }
    for (int k = 0; k < RNG_LEN; k++) {
        uint64_t acc = 0;
        for (int j = 0; j < RNG_LEN; j++) acc += q[j] * s[j];
        y[k] = acc;
        const uint64_t top = q[RNG_LEN - 1];
        memmove(q + 1, q, (RNG_LEN - 1) * sizeof(uint64_t));
        q[0] = top;
        q[RNG_LEN - RNG_TAP] += top;
    }
    for (int k = 0; k < RNG_LEN; k++) {
        const int64_t m = (int64_t)n + k - 606;
        vec[(int)(((334 - m) % RNG_LEN + RNG_LEN) % RNG_LEN)] = y[k];
    }
}

/* srand(seed) of gen_cooked.go: the seeding loop of rng.go with shifts 20 / 10 and no table */
static void gen_srand(int32_t seed, uint64_t *vec) {
    int32_t x = seed;
    for (int i = -20; i < RNG_LEN; i++) {
        x = seedrand(x);
        if (i >= 0) {
            int64_t u = (int64_t)x << 20;
            x = seedrand(x);
            u ^= (int64_t)x << 10;
            x = seedrand(x);
            u ^= (int64_t)x;
            vec[i] = (uint64_t)u;
        }
    }
}

/* test hook: gen_cooked's state after n steps, by direct stepping (jump = 0) or jump-ahead (jump = 1) */
void orc_gorand_gen_state(uint64_t n, int32_t jump, uint64_t *out607) {
    gen_srand(1, out607);
    if (jump) gen_jump(out607, n); else gen_direct(out607, n);
}

static uint64_t g_cooked[RNG_LEN];
static int g_cooked_ready = 0;

const uint64_t *orc_gorand_cooked(void) {
#pragma omp critical(orc_cooked)
    if (!g_cooked_ready) {
        gen_srand(1, g_cooked);
        gen_jump(g_cooked, 7800000000000ull); /* gen_cooked.go: for i := uint64(0); i < 7.8e12; i++ { vrand() } */
        g_cooked_ready = 1;
    }
    return g_cooked;
}

/* rngSource.Seed */
void orc_gorand_seed(orc_gorand *r, int64_t seed) {
    const uint64_t *cooked = orc_gorand_cooked();
    r->tap = 0;
    r->feed = RNG_LEN - RNG_TAP;
    seed = seed % INT32_MAX_;
    if (seed < 0) seed += INT32_MAX_;
    if (seed == 0) seed = 89482311;
    int32_t x = (int32_t)seed;
    for (int i = -20; i < RNG_LEN; i++) {
        x = seedrand(x);
        if (i >= 0) {
            int64_t u = (int64_t)x << 40;
            x = seedrand(x);
            u ^= (int64_t)x << 20;
            x = seedrand(x);
            u ^= (int64_t)x;
            r->vec[i] = (uint64_t)u ^ cooked[i];
        }
    }
}

/* rngSource.Uint64 */
uint64_t orc_gorand_uint64(orc_gorand *r) {
    if (--r->tap < 0) r->tap += RNG_LEN;
    if (--r->feed < 0) r->feed += RNG_LEN;
    const uint64_t x = r->vec[r->feed] + r->vec[r->tap];
    r->vec[r->feed] = x;
    return x;
}

int64_t orc_gorand_int63(orc_gorand *r) { return (int64_t)(orc_gorand_uint64(r) & 0x7fffffffffffffffull); }

static int32_t gorand_int31(orc_gorand *r) { return (int32_t)(orc_gorand_int63(r) >> 32); }

/* Rand.Intn for n <= 2^31 - 1 (Int31n) */
int32_t orc_gorand_intn(orc_gorand *r, int32_t n) {
    if ((n & (n - 1)) == 0) return gorand_int31(r) & (n - 1);
    const int32_t max = (int32_t)((1u << 31) - 1 - (1u << 31) % (uint32_t)n);
    int32_t v = gorand_int31(r);
    while (v > max) v = gorand_int31(r);
    return v % n;
}

/* Rand.Float64 */
double orc_gorand_float64(orc_gorand *r) {
    for (;;) {
        const double f = (double)orc_gorand_int63(r) / 9223372036854775808.0;
        if (f != 1.0) return f;
    }
}

/* ------------------------------------------------------------------ ClutterHandler.Update [V] */

/* the sample angles of clutter.go:45: for a := camHeading - viewFOV/2; a <= camHeading+viewFOV/2; a += viewFOV / 20 */
int32_t orc_clutter_angles(double heading, double fov_degrees, double *out_angles, int32_t cap) {
    const double viewFOV = fov_degrees * 3.14159265358979323846 / 180; /* geom.Radians [R] */
    int32_t n = 0;
    if (!(viewFOV / 20 > 0)) return -1; /* the Go loop would not terminate (or not run) */
    for (double a = heading - viewFOV / 2; a <= heading + viewFOV / 2; a += viewFOV / 20) {
        if (n < cap) out_angles[n] = a;
        n++;
        if (n > 64) break;
    }
    return n;
}

int64_t orc_clutter_update(const uint16_t *level0, int32_t map_w, int32_t map_h, const uint8_t *cell_class,
                           const uint64_t *class_mask, const orc_clutter_def *defs, int32_t n_defs, int64_t map_seed,
                           double clutter_distance, const orc_clutter_query *q, orc_clutter_sprite *out, int64_t cap) {
    int64_t n_out = 0;
    if (n_defs == 0 || clutter_distance <= 0) return 0; /* clutter.go:32-34 */
    int64_t *loaded = NULL; /* keys of spritesByPosition */
    size_t n_loaded = 0, cap_loaded = 0;
    orc_gorand rng;
    for (int32_t ai = 0; ai < q->n_angles; ai++) {
        for (double d = 0.0; d <= clutter_distance; d++) {
            /* geom.LineFromAngle(camX, camY, a, d) [R] */
            const double x2 = q->cam_x + d * q->cs[ai][0];
            const double y2 = q->cam_y + d * q->cs[ai][1];
            const int64_t x = orc_go_int(x2), y = orc_go_int(y2);
            if (x < 0 || y < 0 || x >= (int64_t)map_w || y >= (int64_t)map_h) continue;
            const int64_t posId = (x - 1) * (int64_t)map_w + y;
            if (level0[(size_t)x * map_h + y] > 0) continue; /* IsWallAt(0, x, y), game/model/map.go:886-889 */
            int seen = 0;
            for (size_t k = 0; k < n_loaded && !seen; k++) seen = loaded[k] == posId;
            if (seen) continue;
            if (n_loaded == cap_loaded) {
                cap_loaded = cap_loaded ? cap_loaded * 2 : 256;
                loaded = (int64_t *)realloc(loaded, cap_loaded * sizeof(int64_t));
            }
            loaded[n_loaded++] = posId;
            /* floorPathMatch against FloorTexturePathAt(x, y), resolved per floor-path class by the caller */
            const uint64_t match = class_mask[cell_class[(size_t)x * map_h + y]];
            orc_gorand_seed(&rng, (int64_t)((uint64_t)map_seed + (uint64_t)posId));
            for (int32_t i = 0; i < n_defs; i++) {
                if (!((match >> i) & 1u)) continue;
                if (!(orc_gorand_float64(&rng) <= defs[i].frequency)) continue;
                orc_clutter_sprite s;
                memset(&s, 0, sizeof s);
                s.x = (double)x + orc_gorand_float64(&rng);
                s.y = (double)y + orc_gorand_float64(&rng);
                s.cell_x = (int32_t)x;
                s.cell_y = (int32_t)y;
                s.def = i;
                if (n_out < cap) out[n_out] = s;
                n_out++;
            }
        }
    }
    free(loaded);
    return n_out;
}

/* ------------------------------------------------------------------ Game.getRaycastSprites [V] */

/* model.PointInProximity (game/model/math.go:93-96) */
static int point_in_proximity(double distance, double srcX, double srcY, double tgtX, double tgtY) {
    distance = ceil(distance);
    return fabs(tgtX - srcX) <= distance && fabs(tgtY - srcY) <= distance;
}

/* Game.getRaycastSprites (game/sprites.go:102-145) for one camera: the map sprites that pass
 * `renderDistance < 0 || isTarget || PointInProximity(renderDistance, cam, sprite)` in the order the host walks
 * g.sprites ([D]: sync.Map.Range has no order), then every clutter sprite (:121-124), then the extras the game
 * appends (reticle lead, current nav point, player sprite with a debug camera, :127-142).  out_index: world index,
 * -1-k for clutter k, INT32_MIN+k for extra k.  Returns the count (writes <= cap). */
int32_t orc_raycast_sprites(const orc_sprite *world, int32_t n_world, double cam_x, double cam_y, double render_distance,
                            int32_t target, const orc_sprite *clutter, int32_t n_clutter, const orc_sprite *extras,
                            int32_t n_extras, orc_sprite *out, int32_t *out_index, int32_t cap) {
    int32_t n = 0;
    for (int32_t i = 0; i < n_world; i++) {
        const int doSprite = render_distance < 0 || target == i ||
                             point_in_proximity(render_distance, cam_x, cam_y, world[i].x, world[i].y);
        if (!doSprite) continue;
        if (n < cap) { out[n] = world[i]; out_index[n] = i; }
        n++;
    }
    for (int32_t k = 0; k < n_clutter; k++) {
        if (n < cap) { out[n] = clutter[k]; out_index[n] = -1 - k; }
        n++;
    }
    for (int32_t k = 0; k < n_extras; k++) {
        if (n < cap) { out[n] = extras[k]; out_index[n] = INT32_MIN + k; }
        n++;
    }
    return n;
}
