/* pm3d_oracle_path.c — CPU ORACLE of Pathing.FindPath.  TEST INFRASTRUCTURE ONLY (see pm3d_oracle.h).
 *
 * Restates, literally:
 *   [V] game/model/pathing.go:123-186  Pathing.FindPath: tile lookup by int(pos), astar.Path, reversal into cell
 *       centres, the segmentation into Bezier curves of 7 (+ the pulled-in remainder) steps;
 *   [V] game/model/pathing.go:52-92    Tile.PathNeighbors (order (-1,0) (1,0) (0,-1) (0,1), blockers and the outside
 *       of the map excluded), PathNeighborCost = 1, PathEstimatedCost = Manhattan distance;
 *   [V] game/model/pathing.go:94-112   initPathing: a tile is a blocker iff Level(0)[x][y] != 0;
 *   [V] game/common/bezier/bezier.go   bezier.New (binomial control weights), Curve.Point (Miller's scheme), Curve.Curve;
 *   [R] github.com/harbdog/go-astar v0.0.0-20200827232313-4ecf9e304482 (go.mod:12; a fork of beefsack/go-astar, NOT on
 *       disk): astar.Path with its nodeMap and a container/heap priority queue ordered by `rank <` alone;
 *   [R] Go's container/heap (Init / Push / Pop / Remove / up / down), which decides every tie.
 * On a uniform 4-connected grid with the Manhattan estimate most of the queue holds equal ranks, so WHICH of the many
 * shortest paths comes back is decided by the heap's sift order: up() stops at the first parent that is not larger
 * (`!less(j, i)`), down() prefers the left child unless the right one is strictly smaller and stops when the child
 * is not strictly smaller.  PARITY UNPINNED: neither module can be built here (no Go toolchain).
 *
 * [D] decisions: a start or finish outside the map (a nil *Tile, on which the Go code would panic inside
 * PathNeighbors / never match) is "no path"; costs are small integers, kept as such (the reference's float64
 * arithmetic on them is exact).
 */
#include <stdlib.h>
#include <string.h>

#include "pm3d_oracle.h"

typedef struct path_node {
    uint32_t cost;      /* node.cost */
    uint32_t rank;      /* node.rank */
    int32_t parent;     /* cell index of node.parent, -1 = nil */
    int32_t index;      /* node.index: position in the queue */
    uint8_t open, closed, known; /* known: nodeMap holds it (fresh nodes read cost 0, open = closed = false) */
} path_node;

typedef struct path_pq {
    int32_t *cell;      /* priorityQueue: node pointers, as cell indices */
    int32_t n;
} path_pq;

/* priorityQueue.Less / Swap (go-astar pqueue.go) */
static inline int pq_less(const path_pq *q, const path_node *nodes, int i, int j) {
    return nodes[q->cell[i]].rank < nodes[q->cell[j]].rank;
}
static inline void pq_swap(path_pq *q, path_node *nodes, int i, int j) {
    int32_t t = q->cell[i]; q->cell[i] = q->cell[j]; q->cell[j] = t;
    nodes[q->cell[i]].index = i;
    nodes[q->cell[j]].index = j;
}
/* container/heap up / down */
static void heap_up(path_pq *q, path_node *nodes, int j) {
    for (;;) {
        int i = (j - 1) / 2; /* parent */
        if (i == j || !pq_less(q, nodes, j, i)) break;
        pq_swap(q, nodes, i, j);
        j = i;
    }
}
static int heap_down(path_pq *q, path_node *nodes, int i0, int n) {
    int i = i0;
    for (;;) {
        int j1 = 2 * i + 1;
        if (j1 >= n || j1 < 0) break;
        int j = j1;
        int j2 = j1 + 1;
        if (j2 < n && pq_less(q, nodes, j2, j1)) j = j2;
        if (!pq_less(q, nodes, j, i)) break;
        pq_swap(q, nodes, i, j);
        i = j;
    }
    return i > i0;
}
static void heap_push(path_pq *q, path_node *nodes, int32_t cell) {
    nodes[cell].index = q->n;        /* priorityQueue.Push */
    q->cell[q->n++] = cell;
    heap_up(q, nodes, q->n - 1);
}
static int32_t heap_pop(path_pq *q, path_node *nodes) {
    int n = q->n - 1;
    pq_swap(q, nodes, 0, n);
    heap_down(q, nodes, 0, n);
    int32_t cell = q->cell[--q->n];  /* priorityQueue.Pop */
    nodes[cell].index = -1;
    return cell;
}
static void heap_remove(path_pq *q, path_node *nodes, int i) {
    int n = q->n - 1;
    if (n != i) {
        pq_swap(q, nodes, i, n);
        if (!heap_down(q, nodes, i, n)) heap_up(q, nodes, i);
    }
    int32_t cell = q->cell[--q->n];
    nodes[cell].index = -1;
}

/* astar.Path(from, to) over the level-0 grid; path cells goal -> start into out_cells (<= cap), returns their
 * number, 0 when there is no path.  *expanded = nodes popped. */
static int64_t astar_path(const uint16_t *level0, int32_t map_w, int32_t map_h, int32_t from, int32_t to, path_node *nodes,
                          path_pq *q, int32_t *out_cells, int64_t cap, int64_t *expanded) {
    static const int off[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
    memset(nodes, 0, sizeof(path_node) * (size_t)map_w * map_h);
    q->n = 0;
    nodes[from].known = 1; nodes[from].parent = -1;
    nodes[from].open = 1;
    heap_push(q, nodes, from);
    int64_t popped = 0;
    for (;;) {
        if (q->n == 0) { *expanded = popped; return 0; }
        const int32_t cur = heap_pop(q, nodes);
        popped++;
        nodes[cur].open = 0;
        nodes[cur].closed = 1;
        if (cur == to) {
            int64_t n = 0;
            for (int32_t c = cur; c >= 0; c = nodes[c].parent) {
                if (n < cap) out_cells[n] = c;
                n++;
            }
            *expanded = popped;
            return n;
        }
        const int cx = cur / map_h, cy = cur % map_h;
        for (int k = 0; k < 4; k++) {
            const int nx = cx + off[k][0], ny = cy + off[k][1];
            if (nx < 0 || ny < 0 || nx >= map_w || ny >= map_h) continue;   /* W.Tile() == nil */
            if (level0[(size_t)nx * map_h + ny] != 0) continue;             /* TileKindBlocker */
            const int32_t nb = nx * map_h + ny;
            path_node *nn = &nodes[nb];
            if (!nn->known) { nn->known = 1; nn->parent = -1; }             /* nm.get(): fresh node, cost 0 */
            const uint32_t cost = nodes[cur].cost + 1u;
            if (cost < nn->cost) {
                if (nn->open) heap_remove(q, nodes, nn->index);
                nn->open = 0;
                nn->closed = 0;
            }
            if (!nn->open && !nn->closed) {
                nn->cost = cost;
                nn->open = 1;
                const int tx = to / map_h, ty = to % map_h;
                nn->rank = cost + (uint32_t)(abs(tx - nx) + abs(ty - ny));
                nn->parent = cur;
                heap_push(q, nodes, nb);
            }
        }
    }
}

/* bezier.New(cp).Curve(p) with len(p) = 2 * len(cp) (pathing.go:176-177): n control points, 2n curve points */
static void bezier_curve(const double *cp_xy, int n, double *out_xy) {
    double ctrl[16][2], pt[16][2];
    double w = 0;
    for (int i = 0; i < n; i++) {
        if (i == 0) w = 1;
        else if (i == 1) w = (double)n - 1;
        else w *= (double)(n - i) / (double)i;
        ctrl[i][0] = cp_xy[2 * i] * w;
        ctrl[i][1] = cp_xy[2 * i + 1] * w;
    }
    const int np = 2 * n;
    const double nf = (double)(np - 1);
    for (int k = 0; k < np; k++) {
        const double t = (double)k / nf;
        pt[0][0] = ctrl[0][0]; pt[0][1] = ctrl[0][1];
        double u = t;
        for (int i = 1; i < n; i++) {
            pt[i][0] = ctrl[i][0] * u;
            pt[i][1] = ctrl[i][1] * u;
            u *= t;
        }
        const double t1 = 1 - t;
        double tt = t1;
        double px = pt[n - 1][0], py = pt[n - 1][1];
        for (int i = n - 2; i >= 0; i--) {
            px += pt[i][0] * tt;
            py += pt[i][1] * tt;
            tt *= t1;
        }
        out_xy[2 * k] = px; out_xy[2 * k + 1] = py;
    }
}

/* Pathing.FindPath(startPos, finishPos): curve points into out_xy (<= cap points); returns the number of points,
 * -1 when there is no path; *n_cells = cells of the astar path, *expanded = nodes it popped. */
int64_t orc_find_path(const uint16_t *level0, int32_t map_w, int32_t map_h, double sx, double sy, double fx, double fy,
                      double *out_xy, int64_t cap, int64_t *n_cells, int64_t *expanded) {
    *n_cells = 0; *expanded = 0;
    const int64_t ix = orc_go_int(sx), iy = orc_go_int(sy), jx = orc_go_int(fx), jy = orc_go_int(fy);
    if (ix < 0 || iy < 0 || ix >= map_w || iy >= map_h || jx < 0 || jy < 0 || jx >= map_w || jy >= map_h) return -1;
    const size_t cells = (size_t)map_w * map_h;
    path_node *nodes = (path_node *)malloc(sizeof(path_node) * cells);
    path_pq q;
    q.cell = (int32_t *)malloc(sizeof(int32_t) * cells);
    int32_t *path = (int32_t *)malloc(sizeof(int32_t) * cells);
    const int64_t len = astar_path(level0, map_w, map_h, (int32_t)(ix * map_h + iy), (int32_t)(jx * map_h + jy), nodes, &q,
                                   path, (int64_t)cells, expanded);
    free(nodes); free(q.cell);
    *n_cells = len;
    if (len == 0) { free(path); return -1; }
    /* "astar path returned in reverse order": steps[k] = centre of path[len-1-k] */
    double *steps = (double *)malloc(sizeof(double) * 2 * (size_t)len);
    for (int64_t k = 0; k < len; k++) {
        const int32_t c = path[len - 1 - k];
        steps[2 * k] = (double)(c / map_h) + 0.5;
        steps[2 * k + 1] = (double)(c % map_h) + 0.5;
    }
    free(path);
    int64_t n_out = 0;
#define EMIT(x, y) do { if (n_out < cap) { out_xy[2 * n_out] = (x); out_xy[2 * n_out + 1] = (y); } n_out++; } while (0)
    if (len < 3) {
        for (int64_t k = 0; k < len; k++) EMIT(steps[2 * k], steps[2 * k + 1]);
    } else {
        int64_t segmentIndex = 0, segmentSize = 7, remainingSteps = len;
        while (remainingSteps > 0) {
            if (segmentSize <= 1) {
                EMIT(steps[2 * segmentIndex], steps[2 * segmentIndex + 1]);
                remainingSteps--;
                segmentIndex++;
                continue;
            }
            if (remainingSteps < segmentSize) segmentSize = remainingSteps;
            remainingSteps -= segmentSize - 1;
            if (remainingSteps > 0 && remainingSteps < 3) {
                segmentSize += remainingSteps - 1;
                remainingSteps = 0;
            }
            double seg[32];
            bezier_curve(steps + 2 * segmentIndex, (int)segmentSize, seg);
            for (int64_t k = 0; k < 2 * segmentSize; k++) EMIT(seg[2 * k], seg[2 * k + 1]);
            segmentIndex += segmentSize - 1;
        }
    }
#undef EMIT
    free(steps);
    return n_out;
}

/* batch: queries[n][4] = start x, y, finish x, y; out_xy[n][cap][2]; out_count[n] */
void orc_find_path_batch(const uint16_t *level0, int32_t map_w, int32_t map_h, const double *queries, int64_t n,
                         double *out_xy, int64_t cap, int32_t *out_count, int32_t *out_cells, int64_t *out_expanded,
                         int32_t threads) {
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int64_t i = 0; i < n; i++) {
        int64_t nc = 0, ex = 0;
        const int64_t got = orc_find_path(level0, map_w, map_h, queries[4 * i], queries[4 * i + 1], queries[4 * i + 2],
                                          queries[4 * i + 3], out_xy + (size_t)i * cap * 2, cap, &nc, &ex);
        out_count[i] = (int32_t)got;
        if (out_cells) out_cells[i] = (int32_t)nc;
        if (out_expanded) out_expanded[i] = ex;
    }
}
