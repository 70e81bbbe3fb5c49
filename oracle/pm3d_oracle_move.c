/*
 * pm3d_oracle_move.c — CPU restatement of Game.getValidMove in full (game/collision.go:69-394):
 * walls, the player collider, sprite colliders with their Z overlap, the sorted collision list,
 * the map clamp / level-0 cell test / floor clamp, and the checkAlternate escape heuristics.
 *
 * TEST INFRASTRUCTURE ONLY (see pm3d_oracle.h).  PARITY UNPINNED: the reference has no tests.
 *   [V] follows a file under /root/reference (citation given)
 *   [R] recalled from github.com/harbdog/raycaster-go v1.12.0 package geom / geom3d (go.mod:13; the
 *       module is not on disk): LineCircleIntersection is the Rosetta-code Go routine the module
 *       adopted, Distance2 uses math.Pow(d, 2) (which rounds exactly like d*d), Clamp, Line.Angle,
 *       LineFromAngle, NearlyEqual.
 *   [D] decision of this repo where neither pins the behaviour:
 *       - the reference walks g.sprites (sync.Map.Range: unspecified order) and sorts the collision list
 *         with the unstable sort.Slice; here colliders are visited in list order and ties keep insertion
 *         order, so collisions[0] is the earliest-inserted nearest record;
 *       - a NaN distance never becomes collisions[0];
 *       - the transcendental calls of the checkAlternate branch (atan2/cos/sin/fmod) use the C library,
 *         which can differ from Go's software routines in the last ulp.  That branch runs on the HOST in
 *         the product too (pm3d/collision.py), composed from the batched device call.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pm3d_oracle.h"

#define GEOM_EPS 1e-14 /* [R] geom: const eps = 1e-14 */
#define GEOM_PI 3.14159265358979323846
#define GEOM_PI2 (GEOM_PI * 2)
#define GEOM_HALFPI (GEOM_PI / 2)

static inline double sq(double x) { return x * x; }

/* geom.LineCircleIntersection(li, ci, isSegment) [R]; returns the number of points (0..2) in out[] */
int orc_line_circle_intersection(const double *li, double cx, double cy, double radius, int is_segment,
                                 double *out_xy) {
    int n = 0;
    const double x0 = cx, y0 = cy;
    const double x1 = li[0], y1 = li[1], x2 = li[2], y2 = li[3];
    const double A = y2 - y1;
    const double B = x1 - x2;
    const double C = x2 * y1 - x1 * y2;
    const double a = sq(A) + sq(B);
    double b, c;
    int bnz = 1;
    if (fabs(B) >= GEOM_EPS) {
        b = 2 * (A * C + A * B * y0 - sq(B) * x0);
        c = sq(C) + 2 * B * C * y0 - sq(B) * (sq(radius) - sq(x0) - sq(y0));
    } else {
        b = 2 * (B * C + A * B * x0 - sq(A) * y0);
        c = sq(C) + 2 * A * C * x0 - sq(A) * (sq(radius) - sq(x0) - sq(y0));
        bnz = 0;
    }
    double d = sq(b) - 4 * a * c;
    if (d < 0) return 0;

#define WITHIN(x, y) (fabs(sqrt(sq(x2 - x1) + sq(y2 - y1)) - sqrt(sq((x) - x1) + sq((y) - y1)) - sqrt(sq(x2 - (x)) + sq(y2 - (y)))) < GEOM_EPS)
#define RXY(x, y) do { if (!is_segment || WITHIN(x, y)) { out_xy[n * 2] = (x); out_xy[n * 2 + 1] = (y); n++; } } while (0)
    double x, y;
    if (d == 0) {
        if (bnz) { x = -b / (2 * a); y = -(A * x + C) / B; RXY(x, y); }
        else     { y = -b / (2 * a); x = -(B * y + C) / A; RXY(x, y); }
    } else {
        d = sqrt(d);
        if (bnz) {
            x = (-b + d) / (2 * a); y = -(A * x + C) / B; RXY(x, y);
            x = (-b - d) / (2 * a); y = -(A * x + C) / B; RXY(x, y);
        } else {
            y = (-b + d) / (2 * a); x = -(B * y + C) / A; RXY(x, y);
            y = (-b - d) / (2 * a); x = -(B * y + C) / A; RXY(x, y);
        }
    }
#undef RXY
#undef WITHIN
    return n;
}

/* geom.Distance2 [R]: math.Pow(x2-x1, 2) + math.Pow(y2-y1, 2) */
static inline double distance2(double x1, double y1, double x2, double y2) { return sq(x2 - x1) + sq(y2 - y1); }

/* geom.Clamp [R] */
static inline double geom_clamp(double v, double lo, double hi) {
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}

/* zEntityMinMax (game/collision.go:380-394) [V] */
static inline void z_min_max(double z, double height, int anchor, double *lo, double *hi) {
    *lo = 0; *hi = 0;
    switch (anchor) {
    case 0: *lo = z; *hi = z + height; break;                    /* raycaster.AnchorBottom */
    case 1: *lo = z - height / 2; *hi = z + height / 2; break;   /* AnchorCenter */
    case 2: *lo = z - height; *hi = z; break;                    /* AnchorTop */
    default: break;
    }
}

/* zEntityIntersection (game/collision.go:358-377) [V] */
static inline double z_intersection(double source_z, const orc_mover *src, const orc_collider *tgt) {
    double smin, smax, tmin, tmax;
    z_min_max(source_z, src->height, src->anchor, &smin, &smax);
    z_min_max(tgt->z, tgt->height, tgt->anchor, &tmin, &tmax);
    if (smin > tmax || tmin > smax) return -1;
    const double mid = smin + (smax - smin) / 2;
    return geom_clamp(mid, tmin, tmax);
}

/* model.PointInProximity (game/model/math.go:93-96) [V] */
static inline int in_proximity(double distance, double sx, double sy, double tx, double ty) {
    distance = ceil(distance);
    return fabs(tx - sx) <= distance && fabs(ty - sy) <= distance;
}

typedef struct best_rec { double d2, x, y, z; int32_t entity; int have; } best_rec;

static inline void consider(best_rec *b, double posX, double posY, int32_t entity, double x, double y, double z) {
    const double d2 = distance2(posX, posY, x, y);
    if (!b->have) { b->have = 1; b->d2 = d2; b->x = x; b->y = y; b->z = z; b->entity = entity; return; }
    if (d2 < b->d2 || (b->d2 != b->d2 && d2 == d2)) { b->d2 = d2; b->x = x; b->y = y; b->z = z; b->entity = entity; }
}

/* getValidMove with checkAlternate = false (game/collision.go:69-203, 307-355) [V].
 * pts (optional, cap n_pts_cap): the intersectPoints list in insertion order, for the checkAlternate caller. */
static void valid_move_plain(const orc_scene *s, const double *segs, int32_t n_segs, const orc_collider *col,
                             int32_t n_col, int32_t player, double clip, const orc_mover *m, orc_move_result *r,
                             double *pts, int32_t pts_cap) {
    double posX = m->pos_x, posY = m->pos_y, posZ = m->pos_z;
    memset(r, 0, sizeof(*r));
    r->first_entity = -1;
    r->near_x = NAN; r->near_y = NAN;
    r->x = posX; r->y = posY; r->z = posZ;
    if (m->radius <= 0 || m->height <= 0) return;                                      /* :76-78 */
    if (posX == m->move_x && posY == m->move_y && posZ == m->move_z) return;           /* :79-81 */
    double newX = m->move_x, newY = m->move_y, newZ = m->move_z;
    const double moveLine[4] = {posX, posY, newX, newY};
    /* geom3d.Line3d.Distance [R] */
    const double checkDist = sqrt(sq(newX - posX) + sq(newY - posY) + sq(newZ - posZ));
    int32_t n_points = 0, n_collisions = 0;
    best_rec best = {0, 0, 0, 0, -1, 0};
    /* the point the checkAlternate branch picks at :222-229: nearest to pos, first on ties, never a NaN */
    double near_d2 = INFINITY, near_x = NAN, near_y = NAN;
#define ADD_POINT(px, py) do { if (pts && n_points < pts_cap) { pts[n_points * 2] = (px); pts[n_points * 2 + 1] = (py); } n_points++; \
        const double d2_ = distance2(posX, posY, (px), (py)); if (d2_ < near_d2) { near_d2 = d2_; near_x = (px); near_y = (py); } } while (0)

    /* wall collisions :93-97; the points are kept because the player block below re-reads them */
    int32_t n_wall = 0;
    double *wall_pts = NULL;
    const int player_block = player >= 0 && m->self != player && m->parent != player && !m->destroyed;
    if (player_block) wall_pts = (double *)malloc(sizeof(double) * 2 * (size_t)(n_segs > 0 ? n_segs : 1));
    for (int32_t k = 0; k < n_segs; k++) {
        double px, py;
        if (orc_line_intersection(moveLine, segs + (size_t)k * 4, &px, &py)) {
            if (wall_pts) { wall_pts[n_wall * 2] = px; wall_pts[n_wall * 2 + 1] = py; }
            n_wall++;
            ADD_POINT(px, py);
        }
    }

    /* player collider :100-132 */
    if (player_block) {
        const orc_collider *pl = col + player;
        if (in_proximity(checkDist, newX, newY, pl->x, pl->y)) {
            const double zI = z_intersection(newZ, m, pl);
            double comb[4];
            const int nc = orc_line_circle_intersection(moveLine, pl->x, pl->y, pl->radius + m->radius, 1, comb);
            if (zI >= 0 && nc > 0) {
                /* intersectPoints so far: the wall points, then whatever this loop appends */
                double extra[8];
                int n_extra = 0;
                for (int k = 0; k < nc; k++) {
                    const double chk[4] = {comb[k * 2], comb[k * 2 + 1], pl->x, pl->y};
                    double got[4];
                    const int ng = orc_line_circle_intersection(chk, pl->x, pl->y, pl->radius, 1, got);
                    for (int q = 0; q < ng; q++) {
                        extra[n_extra * 2] = got[q * 2]; extra[n_extra * 2 + 1] = got[q * 2 + 1]; n_extra++;
                        ADD_POINT(got[q * 2], got[q * 2 + 1]);
                    }
                    /* :123-127 appends one collision per point of intersectPoints (ALL of them, as written) */
                    for (int32_t w = 0; w < n_wall; w++) { consider(&best, posX, posY, player, wall_pts[w * 2], wall_pts[w * 2 + 1], zI); n_collisions++; }
                    for (int e = 0; e < n_extra; e++) { consider(&best, posX, posY, player, extra[e * 2], extra[e * 2 + 1], zI); n_collisions++; }
                }
            }
        }
    }
    free(wall_pts);

    /* sprite colliders :135-194 (list order [D]) */
    for (int32_t i = 0; i < n_col; i++) {
        if (i == player) continue;                      /* the player is not in g.sprites */
        const orc_collider *sp = col + i;
        if (m->self == i || m->parent == i || sp->radius <= 0 || sp->destroyed) continue;
        if (!in_proximity(checkDist, newX, newY, sp->x, sp->y)) continue;
        const double zI = z_intersection(newZ, m, sp);
        if (zI < 0) continue;
        const double comb_r = sp->radius + m->radius;
        double comb[4];
        const int nc = orc_line_circle_intersection(moveLine, sp->x, sp->y, comb_r, 1, comb);
        int circle_hit = 0;
        for (int k = 0; k < nc; k++) {
            const double chk[4] = {comb[k * 2], comb[k * 2 + 1], sp->x, sp->y};
            double got[4];
            const int ng = orc_line_circle_intersection(chk, sp->x, sp->y, sp->radius, 1, got);
            for (int q = 0; q < ng; q++) {
                circle_hit = 1;
                ADD_POINT(got[q * 2], got[q * 2 + 1]);
                consider(&best, posX, posY, i, got[q * 2], got[q * 2 + 1], zI);
                n_collisions++;
            }
        }
        if (!circle_hit) {
            const double chkDist = sqrt(distance2(newX, newY, sp->x, sp->y));   /* Line.Distance [R] */
            if (chkDist <= comb_r) {
                ADD_POINT(newX, newY);
                consider(&best, posX, posY, i, newX, newY, zI);
                n_collisions++;
            }
        }
    }
#undef ADD_POINT

    r->n_points = n_points;
    r->n_collisions = n_collisions;
    r->near_x = near_x; r->near_y = near_y;
    if (best.have) { r->first_entity = best.entity; r->first_x = best.x; r->first_y = best.y; r->first_z = best.z; }
    if (n_points > 0) { r->collided = 1; return; }                                      /* :202, :307-311 */

    /* :314-342 */
    int64_t ix = orc_go_int(newX), iy = orc_go_int(newY);
    if (ix < 0 || newX < 0) { newX = clip; ix = 0; }
    else if (ix >= s->map_w) { newX = (double)s->map_w - clip; ix = orc_go_int(newX); }
    if (iy < 0 || newY < 0) { newY = clip; iy = 0; }
    else if (iy >= s->map_h) { newY = (double)s->map_h - clip; iy = orc_go_int(newY); }
    if (ix > s->map_w - 1) ix = s->map_w - 1; /* clip == 0 would index past the edge (Go panics) */
    if (iy > s->map_h - 1) iy = s->map_h - 1;
    if (ix < 0) ix = 0;
    if (iy < 0) iy = 0;
    if (s->grid[ix * s->map_h + iy] <= 0) { posX = newX; posY = newY; }
    else r->collided = 1;
    /* :346-352 */
    posZ = newZ;
    double zMin, zMax;
    z_min_max(0, m->height, m->anchor, &zMin, &zMax);
    zMin = fabs(zMin);
    if (posZ < zMin) { posZ = zMin; r->collided = 1; }
    r->x = posX; r->y = posY; r->z = posZ;
}

void orc_valid_move_batch(const orc_scene *s, const double *segs, int32_t n_segs, const orc_collider *colliders,
                          int32_t n_colliders, int32_t player, double clip, const orc_mover *movers, int64_t n,
                          orc_move_result *out, int32_t threads) {
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
    for (int64_t i = 0; i < n; i++)
        valid_move_plain(s, segs, n_segs, colliders, n_colliders, player, clip, movers + i, out + i, NULL, 0);
}

/* model.AngleDistance (game/model/math.go:149-152) [V] */
static double angle_distance(double a, double b) {
    return fmod(fmod(b - a + GEOM_PI, GEOM_PI2) - GEOM_PI2, GEOM_PI2) + GEOM_PI;
}

/* geom.NearlyEqual [R] */
static int nearly_equal(double a, double b, double epsilon) {
    if (a == b) return 1;
    const double diff = fabs(a - b);
    const double tiny = 4.9406564584124654e-324;
    if (a == 0.0 || b == 0.0 || diff < tiny) return diff < epsilon * tiny;
    return diff / (fabs(a) + fabs(b)) < epsilon;
}

/* getValidMove with checkAlternate (game/collision.go:204-306) [V]: the first result's collision flag and
 * collision list are returned together with the position of the escape attempt. */
void orc_valid_move_alternate(const orc_scene *s, const double *segs, int32_t n_segs, const orc_collider *colliders,
                              int32_t n_colliders, int32_t player, double clip, const orc_mover *m, double heading,
                              double velocity, int check_alternate, orc_move_result *out) {
    const int32_t cap = n_segs + 8 * (n_colliders + 2) + 8;
    double *pts = (double *)malloc(sizeof(double) * 2 * (size_t)cap);
    valid_move_plain(s, segs, n_segs, colliders, n_colliders, player, clip, m, out, pts, cap);
    if (!(out->n_points > 0) || !check_alternate) { free(pts); return; }
    const double posX = m->pos_x, posY = m->pos_y, posZ = m->pos_z;
    double moveX = m->move_x, moveY = m->move_y;
    const double moveZ = m->move_z;
    orc_move_result sub;
    orc_mover t = *m;
#define SUB(mx, my, mz, alt) do { t.move_x = (mx); t.move_y = (my); t.move_z = (mz); \
        orc_valid_move_alternate(s, segs, n_segs, colliders, n_colliders, player, clip, &t, heading, velocity, (alt), &sub); } while (0)
#define TAKE() do { out->x = sub.x; out->y = sub.y; out->z = sub.z; } while (0)
    if (fabs(moveZ - posZ) > 0) {                                                       /* :205-217 */
        SUB(posX, posY, moveZ, 0);
        if (!sub.collided) { TAKE(); }
        else { SUB(moveX, moveY, posZ, 1); TAKE(); }
    } else if (out->n_collisions > 0) {                                                 /* :219-249 */
        double min = INFINITY;
        int32_t minI = -1;
        const int32_t np = out->n_points < cap ? out->n_points : cap;
        for (int32_t i = 0; i < np; i++) {
            const double d2 = distance2(posX, posY, pts[i * 2], pts[i * 2 + 1]);
            if (d2 < min) { min = d2; minI = i; }
        }
        if (minI >= 0) {  /* the reference indexes intersectPoints[-1] (panics) when every distance is NaN/Inf */
            const double intersectAngle = atan2(pts[minI * 2 + 1] - posY, pts[minI * 2] - posX);   /* Line.Angle [R] */
            const double checkDist = sqrt(sq(moveX - posX) + sq(moveY - posY) + sq(moveZ - posZ));
            const double escapeDist = geom_clamp(checkDist, 0, m->radius - 0.001);
            const double escapeAngle = angle_distance(heading, intersectAngle) >= 0 ? intersectAngle - GEOM_HALFPI
                                                                                    : intersectAngle + GEOM_HALFPI;
            /* geom.LineFromAngle [R] */
            const double escapeX = posX + escapeDist * cos(escapeAngle), escapeY = posY + escapeDist * sin(escapeAngle);
            SUB(escapeX, escapeY, moveZ, 0);
            TAKE();
        }
    } else {                                                                            /* :251-305 */
        const double identityX = fabs(0.0 + 1.0 * cos(heading)), identityY = fabs(0.0 + 1.0 * sin(heading));
        if (identityX >= identityY) {
            SUB(moveX, posY, moveZ, 0);
            if (!sub.collided) { TAKE(); }
            else {
                if (nearly_equal(posY, moveY, 0.001)) {
                    double dY = velocity / 2;
                    if (moveY - posY == 0) { if (signbit(moveX)) dY *= -1; }
                    else if (signbit(moveY - posY)) dY *= -1;
                    moveY = posY + dY;
                }
                SUB(posX, moveY, moveZ, 0);
                TAKE();
            }
        } else {
            SUB(posX, moveY, moveZ, 0);
            if (!sub.collided) { TAKE(); }
            else {
                if (nearly_equal(posX, moveX, 0.001)) {
                    double dX = velocity / 2;
                    if (moveX - posX == 0) { if (signbit(moveY)) dX *= -1; }
                    else if (signbit(moveX - posX)) dX *= -1;
                    moveX = posX + dX;
                }
                SUB(moveX, posY, moveZ, 0);
                TAKE();
            }
        }
    }
#undef SUB
#undef TAKE
    free(pts);
}
