// wall_lines.cpp — host mirror of Map.GenerateWallCollisionLines (reference
// game/model/map.go:811-851, with wallLineGenerator :556-809 and createWallLine :854-884).
//
// The Go game keeps calling its own implementation and hands the result to
// pm3d_upload_segments; this mirror exists so synthetic workloads (bench.py, tests) can build
// g.collisionMap without the Go side.  It produces the same segment list in the same order:
// border rectangles first, then one merged outline per wall group in scan order.
//
// Design: instead of a map of per-cell objects, the four directed unit edges of wall cell
// (x, y) live in flat arrays indexed by (x*h + y)*4 + dir; "cancelled" = the neighbour across
// that edge is also a (non-boundary-ring) wall cell, which is exactly when the reference finds
// the opposite directed line in its set.
#include <stdint.h>

#include <vector>

#include "pm3d.h"

namespace {

enum Dir { N = 0, E = 1, S = 2, W = 3 };

struct Edge { double x1, y1, x2, y2; };

inline Edge make_edge(int x, int y, int d) {
    double x1 = x, y1 = y, x2 = x, y2 = y;
    switch (d) {
    case N: y2 += 1; break;                      // bottom-left, going up
    case E: y1 += 1; x2 += 1; y2 += 1; break;    // top-left, going right
    case S: x1 += 1; y1 += 1; x2 += 1; break;    // top-right, going down
    case W: x1 += 1; break;                      // bottom-right, going left
    }
    return {x1, y1, x2, y2};
}

// geom.Rect(x, y, w, h): four edges starting at (x, y)
inline void push_rect(std::vector<Edge> &out, double x, double y, double w, double h) {
    out.push_back({x, y, x, y + h});
    out.push_back({x, y + h, x + w, y + h});
    out.push_back({x + w, y + h, x + w, y});
    out.push_back({x + w, y, x, y});
}

}  // namespace

extern "C" int pm3d_generate_wall_collision_lines(const uint16_t *level0, int w, int h, int has_boundary_wall,
                                                  double clip, double *out_xyxy, int cap) {
    if (!level0 || w < 0 || h < 0 || (cap > 0 && !out_xyxy)) return PM3D_ERR_ARG;
    if (w == 0 || h == 0) return 0;
    std::vector<Edge> lines;
    if (has_boundary_wall) push_rect(lines, 1 + clip, 1 + clip, (double)(w - 2) - 2 * clip, (double)(h - 2) - 2 * clip);
    push_rect(lines, clip, clip, (double)w - 2 * clip, (double)h - 2 * clip);

    auto is_cell = [&](int x, int y) -> bool {
        if (x < 0 || y < 0 || x >= w || y >= h) return false;
        if (has_boundary_wall && (x == 0 || y == 0 || x == w - 1 || y == h - 1)) return false;
        return level0[(size_t)x * h + y] != 0;
    };
    static const int step_x[4] = {0, 1, 0, -1}, step_y[4] = {1, 0, -1, 0};    // cellInDirection
    static const int diag_x[4] = {-1, 1, 1, -1}, diag_y[4] = {1, 1, -1, -1};  // cellInDiagonalFromDirection
    // the neighbour whose opposite edge coincides with edge d of (x, y): N edge is the cell's
    // west border, E edge its north border, S edge its east border, W edge its south border
    static const int opp_x[4] = {-1, 0, 1, 0}, opp_y[4] = {0, 1, 0, -1};

    const size_t cells = (size_t)w * h;
    std::vector<uint8_t> cancelled(cells * 4, 0), visited_line(cells * 4, 0), visited_cell(cells, 0);
    for (int x = 0; x < w; x++)
        for (int y = 0; y < h; y++) {
            if (!is_cell(x, y)) continue;
            for (int d = 0; d < 4; d++)
                if (is_cell(x + opp_x[d], y + opp_y[d])) cancelled[((size_t)x * h + y) * 4 + d] = 1;
        }

    struct Ref { int x, y, d; };
    std::vector<Ref> group;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            if (!is_cell(x, y) || visited_cell[(size_t)x * h + y]) continue;
            const size_t base = ((size_t)x * h + y) * 4;
            if (cancelled[base] && cancelled[base + 1] && cancelled[base + 2] && cancelled[base + 3]) continue;
            group.clear();
            int dir = N;
            const Edge og = make_edge(x, y, N);
            int i = x, j = y, start_dir = dir;
            for (;;) {
                visited_cell[(size_t)i * h + j] = 1;
                const size_t li = ((size_t)i * h + j) * 4 + dir;
                if (cancelled[li] || visited_line[li]) {
                    dir = (dir + 1) & 3;
                    if (dir == start_dir) break;  // "unable to escape cell lines"
                    continue;
                }
                start_dir = dir;
                visited_line[li] = 1;
                group.push_back({i, j, dir});
                // same direction in the adjacent cell
                int pi = i + step_x[dir], pj = j + step_y[dir];
                if (is_cell(pi, pj)) {
                    const size_t pl = ((size_t)pi * h + pj) * 4 + dir;
                    if (!cancelled[pl] && !visited_line[pl]) { i = pi; j = pj; continue; }
                }
                // previous direction in the diagonal cell
                int di = i + diag_x[dir], dj = j + diag_y[dir];
                if (is_cell(di, dj)) {
                    const int dd = (dir + 3) & 3;
                    const size_t dl = ((size_t)di * h + dj) * 4 + dd;
                    if (!cancelled[dl] && !visited_line[dl]) { i = di; j = dj; dir = dd; continue; }
                }
                const Edge e = make_edge(i, j, dir);
                if (e.x2 == og.x1 && e.y2 == og.y1) break;
            }
            // merge runs of the same direction
            int prev = -1;
            for (const Ref &r : group) {
                const Edge e = make_edge(r.x, r.y, r.d);
                if (r.d == prev) { lines.back().x2 = e.x2; lines.back().y2 = e.y2; }
                else lines.push_back(e);
                prev = r.d;
            }
        }
    }
    const int n = (int)lines.size();
    for (int k = 0; k < n && k < cap; k++) {
        out_xyxy[k * 4 + 0] = lines[k].x1; out_xyxy[k * 4 + 1] = lines[k].y1;
        out_xyxy[k * 4 + 2] = lines[k].x2; out_xyxy[k * 4 + 3] = lines[k].y2;
    }
    return n;
}
