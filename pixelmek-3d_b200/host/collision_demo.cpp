// collision_demo.cpp — drives pm3d::Game (game_collision.hpp) from a scenario file (tests/test_host_cpp.py writes it):
//   int32 map_w, map_h, n_segs, n_colliders, n_movers
//   uint16 level0[map_w][map_h]; double segs[n_segs][4]; pm3d_collider colliders[n_colliders] (element 0 = the player)
//   per mover: pm3d_mover (80 B) + double heading, velocity
// and writes, per mover, getValidMove(entity, moveX, moveY, moveZ, checkAlternate=true):
//   double x, y, z; int32 collided, n_collisions, first_entity, _pad; double first_x, first_y, first_z   (64 B)
// then one byte per mover: lineOfSight(entity, colliders[i % n_colliders]).
#include <cstdio>
#include <cstdlib>

#include "game_collision.hpp"

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: collision_demo scenario.bin results.bin\n"); return 2; }
    pm3d_ctx *ctx = nullptr;
    if (pm3d_create(0, nullptr, &ctx) != PM3D_OK) { std::fprintf(stderr, "pm3d_create: %s\n", pm3d_last_error(nullptr)); return 1; }
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 2; }
    int32_t hdr[5];
    if (std::fread(hdr, 4, 5, f) != 5) return 2;
    const int map_w = hdr[0], map_h = hdr[1], n_segs = hdr[2], n_col = hdr[3], n_mov = hdr[4];
    std::vector<uint16_t> level0((size_t)map_w * map_h);
    std::vector<double> segs((size_t)n_segs * 4);
    std::vector<pm3d_collider> col(n_col);
    if (std::fread(level0.data(), 2, level0.size(), f) != level0.size()) return 2;
    if (std::fread(segs.data(), 8, segs.size(), f) != segs.size()) return 2;
    if (std::fread(col.data(), sizeof(pm3d_collider), col.size(), f) != col.size()) return 2;
    try {
        if (pm3d_upload_grid(ctx, level0.data(), 1, map_w, map_h, 1) != PM3D_OK) throw std::runtime_error(pm3d_last_error(ctx));
        pm3d::Game game(ctx, segs);
        std::vector<pm3d::Entity> ents(n_col);
        for (int i = 0; i < n_col; i++) {
            ents[i].x = col[i].x; ents[i].y = col[i].y; ents[i].z = col[i].z; ents[i].collisionRadius = col[i].radius;
            ents[i].collisionHeight = col[i].height; ents[i].anchor = col[i].anchor; ents[i].destroyed = col[i].destroyed != 0;
            ents[i].index = i;
        }
        game.setColliders(&ents[0], std::vector<pm3d::Entity>(ents.begin() + 1, ents.end()));
        FILE *o = std::fopen(argv[2], "wb");
        if (!o) { std::perror(argv[2]); return 2; }
        std::vector<uint8_t> los(n_mov);
        for (int k = 0; k < n_mov; k++) {
            pm3d_mover m;
            double hv[2];
            if (std::fread(&m, sizeof m, 1, f) != 1 || std::fread(hv, 8, 2, f) != 2) return 2;
            pm3d::Entity e;
            e.x = m.pos_x; e.y = m.pos_y; e.z = m.pos_z; e.collisionRadius = m.radius; e.collisionHeight = m.height;
            e.anchor = m.anchor; e.destroyed = m.destroyed != 0; e.heading = hv[0]; e.velocity = hv[1]; e.index = m.self; e.parent = m.parent;
            const pm3d::ValidMove v = game.getValidMove(e, m.move_x, m.move_y, m.move_z, true);
            struct { double x, y, z; int32_t collided, n, first, pad; double fx, fy, fz; } rec = {v.x, v.y, v.z, v.isCollision ? 1 : 0,
                (int32_t)v.collisions.size(), v.collisions.empty() ? -1 : v.collisions[0].entity, 0,
                v.collisions.empty() ? 0.0 : v.collisions[0].x, v.collisions.empty() ? 0.0 : v.collisions[0].y,
                v.collisions.empty() ? 0.0 : v.collisions[0].z};
            std::fwrite(&rec, sizeof rec, 1, o);
            los[k] = game.lineOfSight(&e, &ents[k % n_col]) ? 1 : 0;
        }
        std::fwrite(los.data(), 1, los.size(), o);
        std::fclose(o);
    } catch (const std::exception &ex) {
        std::fprintf(stderr, "collision_demo: %s\n", ex.what());
        return 1;
    }
    std::fclose(f);
    pm3d_destroy(ctx);
    return 0;
}
