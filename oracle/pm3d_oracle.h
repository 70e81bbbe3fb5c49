/*
 * pm3d_oracle.h — CPU ORACLE for the PixelMek 3D raycasting hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (pixelmek-3d_b200/csrc, libpm3d.so) never links, imports or calls anything in oracle/.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the third-party Go module
 * github.com/harbdog/raycaster-go v1.12.0 (reference go.mod:13, go.sum:60-61), which is
 * NOT vendored under /root/reference, and the reference ships no tests or golden
 * vectors.  No Go toolchain exists in this image, so the
 * reference cannot be built or run.  This file restates
 *   [V] what is verifiable on disk (call sites, callback semantics, grid layout,
 *       collision queries) with file:line citations, and
 *   [R] the module's published algorithm (Lode-style DDA raycaster as adapted by
 *       raycaster-go) recalled from the public source, without file:line, and
 *   [D] a design decision of this repo where neither [V] nor [R] pins the behaviour
 *       (chiefly: the integer blit arithmetic Ebitengine performs on the GPU).
 * Every function below carries one of those tags.
 *
 * Numeric contract (mirrors Go on amd64): IEEE-754 binary64, no FMA contraction
 * (compile with -ffp-contract=off), int(float64) truncates toward zero and yields
 * INT64_MIN on NaN/overflow (CVTTSD2SQ), Go `int` is 64-bit and wraps silently.
 */
#ifndef PM3D_ORACLE_H
#define PM3D_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* raycaster.SpriteAnchor [R]: AnchorBottom = iota, AnchorCenter, AnchorTop
 * (used by the reference at game/model/map.go:143-155). */
enum { ORC_ANCHOR_BOTTOM = 0, ORC_ANCHOR_CENTER = 1, ORC_ANCHOR_TOP = 2 };

/* Level grid + texture callbacks, flattened.
 * grid index order is [level][x][y]  (reference game/model/map.go:82, :245-264).   */
typedef struct orc_scene {
    int32_t map_w, map_h;          /* Map.Size()                       map.go:245-251 */
    int32_t stored_levels;         /* len(Map.Levels)                                 */
    int32_t num_levels;            /* Map.NumLevels() = NumRaycastLevels map.go:253   */
    const uint16_t *grid;          /* [stored_levels][map_w][map_h], >0 = wall        */
    const int16_t *wall_lut;       /* [32768][2]: tile for (texIndex, side); -1 = nil.
                                      Encodes TextureAt + MapTexture.GetImage
                                      (texture.go:69-97, map.go:106-115)              */
    const int16_t *floor_tiles;    /* [map_w][map_h] tile per cell, -1 = nil.
                                      Encodes FloorTextureAt (texture.go:119-133)     */
    const uint8_t *atlas;          /* [n_tiles][tex][tex][4] premultiplied RGBA8      */
    int32_t n_tiles, tex_size;     /* resources.TexSize = 256  resources.go:35        */
    int32_t sky_tile, floorbox_tile; /* SetSkyTexture/SetFloorTexture mission.go:78-83 */
    /* sprite sheets: premultiplied RGBA8, row-major, concatenated */
    const uint8_t *sheet_pixels;
    const int64_t *sheet_offset;   /* [n_sheets] byte offset into sheet_pixels        */
    const int32_t *sheet_w, *sheet_h;
    int32_t n_sheets;
    int32_t _pad;
} orc_scene;

/* Camera state the reference sets through K5 setters.
 * Transcendentals (cos/sin/tan) are evaluated by the HOST shim; the hot path sees
 * only dir/plane/pitch_px/cam_z_px so it needs + - * / sqrt only.                    */
typedef struct orc_camera {
    double pos_x, pos_y;           /* SetPosition                    camera.go:31     */
    double cam_z_px;               /* [R] (posZ-0.5)*h               camera.go:32     */
    double dir_x, dir_y;           /* [R] fovDepth*(cos,sin)(heading) camera.go:33    */
    double plane_x, plane_y;       /* [R] perp(dir)*tan(fov/2)        config.go:287   */
    int32_t pitch_px;              /* [R] from SetPitchAngle         camera.go:34     */
    int32_t w, h;                  /* SetViewSize                    config.go:249    */
    int32_t _pad;
    double render_distance;        /* SetRenderDistance, <0 = inf    mission.go:75    */
    double light_falloff;          /* SetLightFalloff                mission.go:90    */
    double global_illumination;    /* SetGlobalIllumination          mission.go:91    */
    uint8_t min_light[4];          /* SetLightRGB                    mission.go:92    */
    uint8_t max_light[4];
} orc_camera;

/* One raycaster.Sprite snapshot (reference game/render/sprites/sprite.go:49-91). */
typedef struct orc_sprite {
    double x, y, z;                /* Pos(), PosZ()                                   */
    double scale;                  /* Scale()                                         */
    double illumination;           /* Illumination()                                  */
    int32_t sheet;                 /* Texture() -> sheet id                           */
    int32_t rx0, ry0, rx1, ry1;    /* TextureRect()                                   */
    uint8_t anchor;                /* VerticalAnchor()                                */
    uint8_t focusable;             /* IsFocusable()                                   */
    uint8_t _pad[2];
} orc_sprite;                      /* 64 bytes */

/* Per (level, column) wall-cast record ("slice") [R]. */
typedef struct orc_column {
    double perp_dist;              /* perpWallDist                                    */
    int64_t line_height;           /* int(h / perpWallDist)                           */
    int64_t draw_start, draw_end;  /* unclamped slice rows [start, end)               */
    int32_t map_x, map_y;          /* hit cell (clamped when the ray left the grid)   */
    int32_t tex_x;                 /* texture column                                  */
    int16_t tile;                  /* atlas tile, -1 = nothing drawn                  */
    uint8_t side;                  /* 0 = x-side, 1 = y-side                          */
    uint8_t hit;                   /* 1 = wall, 2 = left the grid                     */
    uint8_t tint[4];               /* RGB light tint, A = 255                         */
} orc_column;                      /* 56 bytes */

typedef struct orc_rect { int32_t x0, y0, x1, y1; } orc_rect;   /* x0 > x1 = nil */

typedef struct orc_convergence {
    double distance;               /* [D] see pm3d_oracle.c orc_render                */
    double point_x, point_y;
    int32_t sprite;                /* index into the sprite list, -1 = none           */
    int32_t _pad;
} orc_convergence;

/* ---- camera path -------------------------------------------------------------- */

/* K2 wall cast only: columns[level*w + x], zbuf[x] (level 0).  [R] */
void orc_cast_walls(const orc_scene *s, const orc_camera *c, orc_column *columns, double *zbuf);

/* K2+K3+K4: Camera.Update(sprites) followed by Camera.Draw(screen).
 * rgba: [h][w][4] premultiplied RGBA8 out.  rects: [n_sprites] (SetScreenRect).
 * columns/zbuf may be NULL.  threads <= 1 runs single threaded. */
void orc_render(const orc_scene *s, const orc_camera *c, const orc_sprite *sprites, int32_t n_sprites,
                uint8_t *rgba, orc_rect *rects, orc_convergence *conv, orc_column *columns, double *zbuf,
                int32_t threads);

/* far -> near order used for sprite painting; order[i] = sprite index. [R]/[D] */
void orc_sprite_order(const orc_camera *c, const orc_sprite *sprites, int32_t n, int32_t *order);

/* ---- LOS / collision path ------------------------------------------------------ */

/* geom.LineIntersection [R] (call sites game/collision.go:58,94). segs/rays: x1,y1,x2,y2 */
int orc_line_intersection(const double *l1, const double *l2, double *px, double *py);

/* Game.lineOfSight over a batch (game/collision.go:40-65) [V]: out[i] = 1 iff no segment
 * of the collision map intersects ray i. */
void orc_los_batch(const double *segs, int32_t n_segs, const double *rays, int64_t n_rays, uint8_t *out,
                   int32_t threads);

/* Wall part of Game.getValidMove with checkAlternate=false and no sprite colliders
 * (game/collision.go:82-97, 202-203, 307-342) [V]:
 * moves: posX,posY,newX,newY per entry; out_pos: x,y ; out_collided; out_hits = number of
 * intersection points. */
void orc_move_walls_batch(const orc_scene *s, const double *segs, int32_t n_segs, const double *moves,
                          int64_t n_moves, double clip_distance, double *out_pos, uint8_t *out_collided,
                          int32_t *out_hits, int32_t threads);

/* ---- full Game.getValidMove (pm3d_oracle_move.c) -------------------------------- */

/* The moving model.Entity and the requested move (game/collision.go:69-81). */
typedef struct orc_mover {
    double pos_x, pos_y, pos_z;    /* entity.Pos(), entity.PosZ()                     */
    double move_x, move_y, move_z; /* moveX, moveY, moveZ                             */
    double radius, height;         /* CollisionRadius(), CollisionHeight()            */
    int32_t self, parent;          /* collider index of the entity / of entity.Parent(), -1 = none */
    uint8_t anchor, destroyed;     /* Anchor(), IsDestroyed()                         */
    uint8_t _pad[6];
} orc_mover;                       /* 80 bytes */

/* One collidable entity: the player unit or a sprite of a collision type (collision.go:20-29). */
typedef struct orc_collider {
    double x, y, z, radius, height;
    uint8_t anchor, destroyed;
    uint8_t _pad[6];
} orc_collider;                    /* 48 bytes */

typedef struct orc_move_result {
    double x, y, z;                /* returned position                               */
    int32_t collided;              /* isCollision                                     */
    int32_t n_collisions;          /* len(collisionEntities)                          */
    int32_t first_entity;          /* collisions[0].entity (collider index), -1 = none */
    int32_t n_points;              /* len(intersectPoints)                            */
    double first_x, first_y, first_z; /* collisions[0].collision, .collisionZ         */
    double near_x, near_y;         /* intersectPoints[minI] of collision.go:222-229 (NaN when none / all Inf) */
} orc_move_result;                 /* 80 bytes */

/* geom.LineCircleIntersection [R]; writes up to 2 points (x,y pairs), returns their number */
int orc_line_circle_intersection(const double *line, double cx, double cy, double radius, int is_segment,
                                 double *out_xy);
/* getValidMove(entity, x, y, z, checkAlternate=false) for a batch (game/collision.go:69-203, 307-355) [V].
 * player = collider index of g.player.Unit (-1 = none). */
void orc_valid_move_batch(const orc_scene *s, const double *segs, int32_t n_segs, const orc_collider *colliders,
                          int32_t n_colliders, int32_t player, double clip_distance, const orc_mover *movers,
                          int64_t n, orc_move_result *out, int32_t threads);
/* getValidMove with the checkAlternate escape heuristics (game/collision.go:204-306) [V] for one entity;
 * heading = entity.Heading(), velocity = entity.Velocity(). */
void orc_valid_move_alternate(const orc_scene *s, const double *segs, int32_t n_segs, const orc_collider *colliders,
                              int32_t n_colliders, int32_t player, double clip_distance, const orc_mover *m,
                              double heading, double velocity, int check_alternate, orc_move_result *out);

/* ---- presentation step after the raycast (pm3d_oracle_present.c) ----------------- */

/* destination -> source index of the 1/renderScale nearest scale (game/scene_game.go:170-177) [D] */
void orc_present_axis_map(int32_t n_out, int32_t n_src, double render_scale, int32_t *map);
/* rayScreen (sw x sh RGBA8 premultiplied) -> overlayScreen (ow x oh): scale, then copy (mode 0) or the CRT
 * shader without / with curvature (mode 1 / 2; game/resources/shaders/crt.kage, game/scene_game.go:179-185) */
void orc_present(const uint8_t *src, int32_t sw, int32_t sh, int32_t ow, int32_t oh, double render_scale, int32_t mode,
                 uint8_t *out, int32_t threads);

/* ---- Go math/rand + ClutterHandler.Update (pm3d_oracle_clutter.c) ------------------ */

/* rngSource of Go's math/rand [R]; pinned by known answers (tests/test_clutter_cpu.py) */
typedef struct orc_gorand {
    int32_t tap, feed;
    uint64_t vec[607];
} orc_gorand;
const uint64_t *orc_gorand_cooked(void);                /* the rngCooked table, regenerated */
void orc_gorand_gen_state(uint64_t n, int32_t jump, uint64_t *out607); /* gen_cooked.go state after n steps */
void orc_gorand_seed(orc_gorand *r, int64_t seed);      /* Rand.Seed / rand.NewSource */
uint64_t orc_gorand_uint64(orc_gorand *r);
int64_t orc_gorand_int63(orc_gorand *r);
int32_t orc_gorand_intn(orc_gorand *r, int32_t n);
double orc_gorand_float64(orc_gorand *r);

/* one model.MapClutter entry (game/model/map.go:128-133), resolved by the host */
typedef struct orc_clutter_def {
    double frequency;              /* Frequency                                       */
    double scale;                  /* Height / model.METERS_PER_UNIT (clutter.go:88)  */
    int32_t sheet;                 /* g.tex.TextureImage(clutter.Image)               */
    int32_t rx0, ry0, rx1, ry1;    /* GetSpriteSheetSlices(img, 1, 1) rect            */
    int32_t _pad;
} orc_clutter_def;                 /* 40 bytes */
/* one camera: position and (cos, sin) of the sample angles of clutter.go:45 */
typedef struct orc_clutter_query {
    double cam_x, cam_y;
    int32_t n_angles, _pad;
    double cs[21][2];
} orc_clutter_query;               /* 360 bytes */
typedef struct orc_clutter_sprite {
    double x, y;                   /* BasicVisualEntity position (z = 0, AnchorBottom) */
    int32_t cell_x, cell_y;        /* the cell whose posId seeds the generator        */
    int32_t def, _pad;             /* index into Map.Clutter                          */
} orc_clutter_sprite;              /* 32 bytes */
/* the sample angles of clutter.go:44-45; returns their number (<= 21 by construction), -1 for fov <= 0 */
int32_t orc_clutter_angles(double heading, double fov_degrees, double *out_angles, int32_t cap);
/* ClutterHandler.Update(g, forceUpdate=true) on an empty handler (game/clutter.go:25-117) [V]: the clutter
 * sprites in view of one camera, in creation order.  level0 / cell_class: [map_w][map_h]; class_mask[c] bit i =
 * clutter i's FloorPathMatch accepts the floor texture path of class c.  Returns the count (writes <= cap). */
int64_t orc_clutter_update(const uint16_t *level0, int32_t map_w, int32_t map_h, const uint8_t *cell_class,
                           const uint64_t *class_mask, const orc_clutter_def *defs, int32_t n_defs, int64_t map_seed,
                           double clutter_distance, const orc_clutter_query *q, orc_clutter_sprite *out, int64_t cap);

/* Game.getRaycastSprites (game/sprites.go:102-145) for one camera [V]: map sprites passing the proximity test in
 * world order, then clutter, then extras.  out_index: world index, -1-k for clutter k, INT32_MIN+k for extra k. */
int32_t orc_raycast_sprites(const orc_sprite *world, int32_t n_world, double cam_x, double cam_y, double render_distance,
                            int32_t target, const orc_sprite *clutter, int32_t n_clutter, const orc_sprite *extras,
                            int32_t n_extras, orc_sprite *out, int32_t *out_index, int32_t cap);

/* Go int(float64) on amd64. */
int64_t orc_go_int(double v);

/* Pathing.FindPath (game/model/pathing.go:123-186) [V] over astar.Path of go-astar [R] and container/heap [R]
 * (pm3d_oracle_path.c): the smoothed path from start to finish over the level-0 grid [map_w][map_h]; returns the
 * number of curve points (writes <= cap into out_xy), -1 when there is no path. */
int64_t orc_find_path(const uint16_t *level0, int32_t map_w, int32_t map_h, double sx, double sy, double fx, double fy,
                      double *out_xy, int64_t cap, int64_t *n_cells, int64_t *expanded);
void orc_find_path_batch(const uint16_t *level0, int32_t map_w, int32_t map_h, const double *queries, int64_t n,
                         double *out_xy, int64_t cap, int32_t *out_count, int32_t *out_cells, int64_t *out_expanded,
                         int32_t threads);

#ifdef __cplusplus
}
#endif
#endif
