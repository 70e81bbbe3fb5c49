// game_collision.hpp — header-only C++ mirror of the reference's collision boundary above the C-ABI
// (SURVEY.md §8a rows L1 / L2, §8b): same names and argument meaning as the Go methods
//
//   (*Game).lineOfSight(source, target)                       game/collision.go:40-65
//   (*Game).getValidMove(entity, x, y, z, checkAlternate)     game/collision.go:69-355
//
// Every geometric test runs on the device (pm3d_los_batch / pm3d_valid_move_batch).  What lives here is what the
// reference keeps in scalar control flow ON TOP of those tests: the checkAlternate escape heuristics
// (collision.go:204-306), which call getValidMove again with modified targets and use the entity heading through
// atan2 / cos / sin (host transcendentals, as in raycaster_camera.hpp).  The Go host keeps its own Game; this class
// is what a C++ host would write, and what the tests drive (tests/test_host_cpp.py).
#pragma once
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "pm3d.h"

namespace pm3d {

constexpr double kClipDistance = 0.01;  // game/game.go:33

// the subset of model.Entity (game/model/entity.go:12-50) getValidMove reads
struct Entity {
    double x = 0, y = 0, z = 0;
    double collisionRadius = 0, collisionHeight = 0;
    int anchor = PM3D_ANCHOR_BOTTOM;
    bool destroyed = false;
    double heading = 0, velocity = 0;
    int index = -1;   // position in the collider list (set by Game::setColliders), -1 = not a collider
    int parent = -1;  // collider index of Parent(), -1 = none
};

// game/collision.go:14-18
struct EntityCollision {
    int entity = -1;  // collider index
    double x = 0, y = 0, z = 0;
};

struct ValidMove {
    double x = 0, y = 0, z = 0;
    bool isCollision = false;
    std::vector<EntityCollision> collisions;  // collisions[0] when there is one (all the callers read)
};

class Game {
public:
    // collisionMap: g.collisionMap as x1, y1, x2, y2 per segment (Map.GenerateWallCollisionLines)
    Game(pm3d_ctx *ctx, const std::vector<double> &collisionMap) : ctx_(ctx) {
        check(pm3d_upload_segments(ctx_, collisionMap.data(), (int)(collisionMap.size() / 4)), "pm3d_upload_segments");
        setColliders(nullptr, {});
    }

    // g.player.Unit and the sprites of the collision types for this tick (collision.go:20-29)
    void setColliders(const Entity *player, const std::vector<Entity> &sprites) {
        std::vector<pm3d_collider> c;
        auto push = [&](const Entity &e) {
            pm3d_collider r;
            std::memset(&r, 0, sizeof r);
            r.x = e.x; r.y = e.y; r.z = e.z; r.radius = e.collisionRadius; r.height = e.collisionHeight;
            r.anchor = (uint8_t)e.anchor; r.destroyed = e.destroyed ? 1 : 0;
            c.push_back(r);
        };
        if (player) push(*player);
        for (const Entity &e : sprites) push(e);
        check(pm3d_upload_colliders(ctx_, c.data(), (int)c.size(), player ? 0 : -1), "pm3d_upload_colliders");
    }

    bool lineOfSight(const Entity *source, const Entity *target) {
        if (!source || !target) return false;
        const double ray[4] = {source->x, source->y, target->x, target->y};
        uint8_t out = 0;
        check(pm3d_los_batch(ctx_, ray, 1, &out, 0), "pm3d_los_batch");
        return out != 0;
    }

    ValidMove getValidMove(const Entity &e, double moveX, double moveY, double moveZ, bool checkAlternate) {
        pm3d_mover m;
        std::memset(&m, 0, sizeof m);
        m.pos_x = e.x; m.pos_y = e.y; m.pos_z = e.z; m.move_x = moveX; m.move_y = moveY; m.move_z = moveZ;
        m.radius = e.collisionRadius; m.height = e.collisionHeight; m.self = e.index; m.parent = e.parent;
        m.anchor = (uint8_t)e.anchor; m.destroyed = e.destroyed ? 1 : 0;
        pm3d_move_result r;
        check(pm3d_valid_move_batch(ctx_, &m, 1, kClipDistance, &r, 0), "pm3d_valid_move_batch");
        ValidMove v;
        v.x = r.x; v.y = r.y; v.z = r.z; v.isCollision = r.collided != 0;
        if (r.n_collisions > 0) v.collisions.push_back({r.first_entity, r.first_x, r.first_y, r.first_z});
        if (!(r.n_points > 0 && checkAlternate)) return v;

        const double posX = e.x, posY = e.y, posZ = e.z;
        auto moved = [&](const ValidMove &n) { v.x = n.x; v.y = n.y; v.z = n.z; return v; };
        if (std::fabs(moveZ - posZ) > 0) {  // collision.go:205-217: try the vertical move alone, else the horizontal one
            const ValidMove zOnly = getValidMove(e, posX, posY, moveZ, false);
            if (!zOnly.isCollision) return moved(zOnly);
            return moved(getValidMove(e, moveX, moveY, posZ, true));
        }
        if (!v.collisions.empty()) {  // :219-249: slide away from the closest intersection point
            const double nx = r.near_x, ny = r.near_y;
            if (std::isnan(nx)) return v;
            const double intersectAngle = std::atan2(ny - posY, nx - posX);
            const double dx = moveX - posX, dy = moveY - posY, dz = moveZ - posZ;
            const double checkDist = std::sqrt(dx * dx + dy * dy + dz * dz);
            const double hi = e.collisionRadius - 0.001;
            const double escapeDist = checkDist < 0 ? 0.0 : (checkDist > hi ? hi : checkDist);  // geom.Clamp
            const double escapeAngle = angleDistance(e.heading, intersectAngle) >= 0 ? intersectAngle - kHalfPi : intersectAngle + kHalfPi;
            const double escapeX = posX + escapeDist * std::cos(escapeAngle), escapeY = posY + escapeDist * std::sin(escapeAngle);
            return moved(getValidMove(e, escapeX, escapeY, moveZ, false));
        }
        const double identityX = std::fabs(0.0 + 1.0 * std::cos(e.heading)), identityY = std::fabs(0.0 + 1.0 * std::sin(e.heading));
        if (identityX >= identityY) {  // :255-279: the X move alone, else nudge along Y
            const ValidMove xOnly = getValidMove(e, moveX, posY, moveZ, false);
            if (!xOnly.isCollision) return moved(xOnly);
            if (nearlyEqual(posY, moveY, 0.001)) {
                double dY = e.velocity / 2;
                if (moveY - posY == 0) { if (std::signbit(moveX)) dY *= -1; }
                else if (std::signbit(moveY - posY)) dY *= -1;
                moveY = posY + dY;
            }
            return moved(getValidMove(e, posX, moveY, moveZ, false));
        }
        const ValidMove yOnly = getValidMove(e, posX, moveY, moveZ, false);  // :280-304
        if (!yOnly.isCollision) return moved(yOnly);
        if (nearlyEqual(posX, moveX, 0.001)) {
            double dX = e.velocity / 2;
            if (moveX - posX == 0) { if (std::signbit(moveY)) dX *= -1; }
            else if (std::signbit(moveX - posX)) dX *= -1;
            moveX = posX + dX;
        }
        return moved(getValidMove(e, moveX, posY, moveZ, false));
    }

private:
    static constexpr double kPi = 3.14159265358979323846, kHalfPi = kPi / 2, kPi2 = kPi * 2;
    // model.AngleDistance (game/model/math.go:149-152)
    static double angleDistance(double a, double b) { return std::fmod(std::fmod(b - a + kPi, kPi2) - kPi2, kPi2) + kPi; }
    // geom.NearlyEqual
    static bool nearlyEqual(double a, double b, double epsilon) {
        if (a == b) return true;
        const double diff = std::fabs(a - b), tiny = 5e-324;
        if (a == 0.0 || b == 0.0 || diff < tiny) return diff < epsilon * tiny;
        return diff / (std::fabs(a) + std::fabs(b)) < epsilon;
    }
    void check(int rc, const char *what) {
        if (rc != PM3D_OK) throw std::runtime_error(std::string(what) + ": " + pm3d_last_error(ctx_));
    }
    pm3d_ctx *ctx_;
};

}  // namespace pm3d
