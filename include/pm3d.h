/*
 * pm3d.h — C-ABI of libpm3d.so: the B200-native (sm_100a CUDA) raycasting hot path of
 * PixelMek 3D behind the interfaces the game's Go code drives.
 *
 * Each entry point names the reference interface it replaces; paths are relative to the
 * reference repo (pixelmek-3d/pixelmek-3d).  The arithmetic the camera entry points replace
 * lives in the un-vendored module github.com/harbdog/raycaster-go v1.12.0 (go.mod:13), so
 * those citations are the reference's CALL SITES into it.  INTEGRATION.md shows the cgo
 * binding a maintainer adds on the Go side.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.  Every function returns 0 on
 *     success and a negative pm3d_status otherwise; pm3d_last_error() has the message.
 *     There is NO CPU fallback: without a CUDA device every compute call fails.
 *   - grid order is [level][x][y] exactly like model.Map.Levels (game/model/map.go:82).
 *   - all RGBA8 images are premultiplied alpha, row-major (what ebiten.Image / image.RGBA hold).
 *   - transcendental functions stay on the host: a view carries dir/plane/pitch already
 *     evaluated (Go's math.Sin/Cos are software routines that differ from CUDA's in the
 *     last ulp); the device path uses only + - * / sqrt in IEEE binary64 without FMA
 *     contraction so hit cells, sides and texture columns are bit-exact.
 *   - a ctx is single-threaded and behaves like ONE CUDA stream towards its caller: calls take effect in the
 *     order they are made.  Internally frames run on further streams (the cast/setup launch of a frame beside
 *     the compose launch of the previous one; PM3D_OPT_LANES); the library orders those streams against the ctx
 *     stream itself — a frame call waits for everything enqueued before it (sprite lists, clutter, copies), and
 *     every later non-frame call waits for the frames enqueued before it.  Only frames dealt to different LANES
 *     are unordered among themselves.  Use one ctx per GPU.
 */
#ifndef PM3D_H
#define PM3D_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM3D_VERSION 100

typedef struct pm3d_ctx pm3d_ctx;

typedef enum pm3d_status {
    PM3D_OK = 0,
    PM3D_ERR_ARG = -1,      /* bad argument / call order */
    PM3D_ERR_CUDA = -2,     /* CUDA runtime failure (incl. "no device": no CPU fallback) */
    PM3D_ERR_CAPACITY = -3, /* exceeds a limit fixed by pm3d_set_view */
    PM3D_ERR_NCCL = -4      /* pm3d_comm_* / pm3d_group_*: libnccl missing, no communicator, or an NCCL failure */
} pm3d_status;

/* raycaster.SpriteAnchor as used at game/model/map.go:143-155 */
enum { PM3D_ANCHOR_BOTTOM = 0, PM3D_ANCHOR_CENTER = 1, PM3D_ANCHOR_TOP = 2 };

/* memory-space flags for the batched calls */
enum {
    PM3D_OUT_DEVICE = 1u << 0, /* output pointers are device memory (no D2H copy) */
    PM3D_IN_DEVICE = 1u << 1,  /* input arrays are device memory (no H2D copy)  */
    PM3D_NO_SYNC = 1u << 2     /* enqueue only; caller syncs with pm3d_sync     */
};

/* One camera view = the state set through the raycaster.Camera setters:
 *   SetPosition/SetPositionZ/SetHeadingAngle/SetPitchAngle  game/camera.go:31-34
 *   SetFovAngle game/config.go:290, game/input.go:767,771
 *   SetRenderDistance game/mission.go:75; SetLightFalloff/SetGlobalIllumination/SetLightRGB
 *   game/mission.go:90-92.   (host mirror: pixelmek-3d_b200/host/raycaster_camera.hpp)     */
typedef struct pm3d_view {
    double pos_x, pos_y;        /* camera position in cells                          */
    double cam_z_px;            /* (posZ - 0.5) * h, camera height in pixel units    */
    double dir_x, dir_y;        /* fovDepth * (cos, sin)(heading)                    */
    double plane_x, plane_y;    /* camera plane, |plane| = |dir| * tan(fov/2)        */
    int32_t pitch_px;           /* pitch as a pixel offset of the horizon            */
    int32_t _reserved[3];
    double render_distance;     /* cells; < 0 = infinite (game/sprites.go:113)       */
    double light_falloff;
    double global_illumination;
    uint8_t min_light[4];       /* RGB, [3] ignored                                  */
    uint8_t max_light[4];
} pm3d_view;                    /* 104 bytes */

/* One raycaster.Sprite snapshot (game/render/sprites/sprite.go:49-91). */
typedef struct pm3d_sprite {
    double x, y, z;             /* Pos(), PosZ()                                     */
    double scale;               /* Scale()                                           */
    double illumination;        /* Illumination()                                    */
    int32_t sheet;              /* Texture(): id given to pm3d_upload_sprite_sheets  */
    int32_t rx0, ry0, rx1, ry1; /* TextureRect() inside that sheet                   */
    uint8_t anchor;             /* VerticalAnchor()                                  */
    uint8_t focusable;          /* IsFocusable()                                     */
    uint8_t _pad[2];
} pm3d_sprite;                  /* 64 bytes */

/* Sprite.SetScreenRect argument (sprite.go:85); x0 > x1 means nil. */
typedef struct pm3d_rect { int32_t x0, y0, x1, y1; } pm3d_rect;

/* Camera.GetConvergenceSprite / convergence point (game/scene_game.go:164, game/game.go:576-629) */
typedef struct pm3d_convergence {
    double distance;
    double point_x, point_y;
    int32_t sprite;             /* index into that view's sprite list, -1 = none     */
    int32_t _pad;
} pm3d_convergence;

/* Per (level, column) wall-cast record, for parity checks (pm3d_download_columns). */
typedef struct pm3d_column {
    double perp_dist;
    int64_t line_height;
    int64_t draw_start, draw_end;
    int32_t map_x, map_y;       /* hit cell                                          */
    int32_t tex_x;              /* texture column                                    */
    int16_t tile;               /* atlas tile drawn, -1 = none                       */
    uint8_t side;               /* 0 = x side, 1 = y side                            */
    uint8_t hit;                /* 1 = wall, 2 = ray left the grid                   */
    uint8_t tint[4];
} pm3d_column;                  /* 56 bytes */

/* ------------------------------------------------------------------ lifetime */

/* device: CUDA ordinal.  stream: a cudaStream_t to enqueue on, or NULL for a private
 * non-blocking stream. */
int pm3d_create(int device, void *stream, pm3d_ctx **out);
void pm3d_destroy(pm3d_ctx *ctx);
/* message of the last failure on ctx (ctx may be NULL for pm3d_create failures) */
const char *pm3d_last_error(const pm3d_ctx *ctx);
int pm3d_version(void);
int pm3d_device_count(void);
int pm3d_sync(pm3d_ctx *ctx);

/* ------------------------------------------------------------------ mission-time uploads
 * raycaster.NewCamera(w, h, texSize, Map, TextureHandler) — game/mission.go:74 — resolved
 * once per mission instead of called back per ray/pixel.                                 */

/* raycaster.Map: NumLevels()/Level(l) (game/model/map.go:253-264).  Levels above
 * stored_levels repeat the top stored level, as Level() does. */
int pm3d_upload_grid(pm3d_ctx *ctx, const uint16_t *grid, int stored_levels, int map_w, int map_h,
                     int num_raycast_levels);
/* TextureHandler.TextureAt + MapTexture.GetImage (game/texture/texture.go:69-97,
 * game/model/map.go:106-115): for texture index tex_index[i] the tile drawn on side 0 / 1
 * (-1 = nil image).  Unlisted indices draw nothing. */
int pm3d_upload_wall_textures(pm3d_ctx *ctx, const int32_t *tex_index, const int32_t *tile_side0,
                              const int32_t *tile_side1, int n);
/* the tex_size x tex_size wall/floor/sky images (resources.TexSize = 256); tex_size must be
 * a power of two */
int pm3d_upload_atlas(pm3d_ctx *ctx, const uint8_t *rgba, int n_tiles, int tex_size);
/* TextureHandler.FloorTextureAt (game/texture/texture.go:119-133): tile per cell [x][y],
 * -1 = nil (floor texturing disabled) */
int pm3d_upload_floor_tiles(pm3d_ctx *ctx, const int16_t *tiles, int map_w, int map_h);
/* Camera.SetSkyTexture / SetFloorTexture (game/mission.go:78-83); -1 = none */
int pm3d_set_box_tiles(pm3d_ctx *ctx, int sky_tile, int floorbox_tile);
/* Sprite.Texture() images (sheets are uploaded once; per frame a sprite names sheet + rect,
 * game/render/sprites/sprite.go:161-185) */
int pm3d_upload_sprite_sheets(pm3d_ctx *ctx, const uint8_t *pixels, const int64_t *byte_offset,
                              const int32_t *width, const int32_t *height, int n_sheets);
/* Camera.SetViewSize (game/config.go:249) + capacity for batched calls */
int pm3d_set_view(pm3d_ctx *ctx, int width, int height, int max_views, int max_sprites_total);
/* options: PM3D_OPT_KEEP_COLUMNS (default 0) keeps the pm3d_column records of every view for
 * pm3d_download_columns; PM3D_OPT_FORCE_PAINTER (default 0) disables the front-to-back
 * shortcut of the compose kernel (A/B check: output must not change); PM3D_OPT_PIPELINE = n
 * (1..4, default 1) keeps n internal frame slots so that, for PM3D_NO_SYNC renders into HOST
 * memory, the device-to-host copy of frame i (on a second stream) overlaps the kernels of frame
 * i+1; results are valid after pm3d_sync (the rectangles and convergence travel through mapped
 * pinned memory, not the copy engine, and are handed to out_rects / out_conv by pm3d_sync).  Call it after pm3d_set_view.  PM3D_OPT_OVERLAP (default 1): the
 * per-frame intermediate buffers exist twice and the cast/setup launch of frame i+1 runs on a second stream
 * beside the compose launch of frame i (enqueue-only callers: pm3d_render_resident, PM3D_NO_SYNC); 0 = one
 * set of buffers, one stream.
 * PM3D_OPT_LANES = n (1..4, default 1; needs PM3D_OPT_OVERLAP = 1): pm3d_render_resident deals successive calls
 * round-robin to n lanes, each with its own streams and intermediate buffers, so INDEPENDENT frame batches (different
 * cameras, or frames whose results are only read after pm3d_sync) run concurrently: a single 1080p frame is two
 * launches that leave most of the chip idle (24 CTAs of serial ray walks, then a compose launch whose tail is a few
 * sprite-heavy warps), three to four lanes double the frames per second.  Calls on one lane stay ordered; calls on
 * different lanes are unordered among themselves, so two in-flight calls must not write the same output buffer.
 * Results of "the last render" (pm3d_download_results, pm3d_frame_checksums, pm3d_present with d_src NULL) wait for
 * that call's lane; pm3d_timer_* bracket all lanes.  pm3d_render (host buffers) always uses lane 0. */
/* PM3D_OPT_DDA_JUMP (default 1): the wall cast advances between two probes of the map in closed form (a certified
 * state of the reference's DDA after N steps) instead of stepping; 0 runs the sequential loop only (A/B check:
 * output must not change). */
enum { PM3D_OPT_KEEP_COLUMNS = 1, PM3D_OPT_FORCE_PAINTER = 2, PM3D_OPT_PIPELINE = 3, PM3D_OPT_OVERLAP = 4, PM3D_OPT_LANES = 5,
       PM3D_OPT_DDA_JUMP = 6 };
int pm3d_set_option(pm3d_ctx *ctx, int option, int value);

/* ------------------------------------------------------------------ frame path
 * Camera.Update(sprites) + GetConvergenceSprite() + Camera.Draw(screen)
 * (game/scene_game.go:161-167) for n_views independent cameras in one call.
 *   sprites: all views' sprite snapshots concatenated; sprite_offset[n_views+1] delimits them
 *   out_rgba: n_views * h * w * 4 bytes, premultiplied RGBA8 (NULL = keep on device only)
 *   out_rects: one per sprite (SetScreenRect); out_conv: one per view.  Either may be NULL.
 *   With PM3D_NO_SYNC every host output (out_rgba, out_rects, out_conv) should be pinned memory
 *   (pm3d_host_alloc): a copy into pageable memory blocks the caller until the stream reaches it.
 *   flags: PM3D_OUT_DEVICE -> out_rgba is device memory.                                  */
int pm3d_render(pm3d_ctx *ctx, const pm3d_view *views, int n_views, const pm3d_sprite *sprites,
                const int32_t *sprite_offset, uint8_t *out_rgba, pm3d_rect *out_rects,
                pm3d_convergence *out_conv, uint32_t flags);
/* Same frame path with every argument already resident in HBM (device pointers) and nothing
 * copied or synchronised: the call only enqueues the kernels, ordered after everything enqueued on
 * the ctx before it (e.g. the PM3D_NO_SYNC pm3d_raycast_sprites_batch that writes d_sprites /
 * d_sprite_offset) — see "Conventions".  Per-sprite rectangles and per-view convergence stay on the
 * device until pm3d_download_results.
 * n_sprites_total = d_sprite_offset[n_views], or an upper bound of it (e.g. the max_total given to
 * pm3d_raycast_sprites_batch) when the host has not read the total.  The offsets live in HBM, so the
 * host cannot validate them: the view setup does, on the device — a view whose entries are negative,
 * not monotone, list more than 2048 sprites or reach past n_sprites_total is rendered WITHOUT sprites
 * and the next pm3d_sync / pm3d_download_results returns PM3D_ERR_CAPACITY naming the view. */
int pm3d_render_resident(pm3d_ctx *ctx, const pm3d_view *d_views, int n_views, const pm3d_sprite *d_sprites,
                         const int32_t *d_sprite_offset, int n_sprites_total, uint8_t *d_out_rgba);
/* ---- recorded frame sequences.  A headless server that renders the SAME sequence of resident frames again and again
 * (a ring of unit cameras whose poses it rewrites in place in device memory, a replay) pays ~8 runtime calls per
 * frame to enqueue it.  Between pm3d_frames_capture_begin and pm3d_frames_capture_end the pm3d_render_resident calls
 * (and nothing else) are not executed but captured — from every stream and lane the frames would run on — into one
 * CUDA graph; pm3d_frames_replay enqueues all of them again with ONE launch on the ctx stream: same device pointers,
 * whatever they hold at that time, same kernels, same dependencies.  Ordering: a replay follows everything enqueued
 * on the ctx before it and precedes everything after it, like any other call.  PM3D_OPT_LANES / PM3D_OPT_OVERLAP /
 * pm3d_set_view must not change between capture and replay (the graph holds the buffers of that configuration). */
int pm3d_frames_capture_begin(pm3d_ctx *ctx);
int pm3d_frames_capture_end(pm3d_ctx *ctx, int *out_graph);
int pm3d_frames_replay(pm3d_ctx *ctx, int graph);
int pm3d_frames_graph_destroy(pm3d_ctx *ctx, int graph);

/* rects (n_sprites of the last render) and convergence (n_views) of the last render; either may be NULL */
int pm3d_download_results(pm3d_ctx *ctx, pm3d_rect *out_rects, pm3d_convergence *out_conv);
/* wall-cast records + z-buffer of one view of the last pm3d_render (parity/debug).
 * columns: [num_levels][w]; zbuf: [w] */
int pm3d_download_columns(pm3d_ctx *ctx, int view, pm3d_column *columns, double *zbuf);
/* 64-bit FNV-1a checksum of each view's framebuffer of the last pm3d_render, computed on the
 * device (compact per-view result for sharded batches). */
int pm3d_frame_checksums(pm3d_ctx *ctx, int n_views, uint64_t *out, uint32_t flags);
/* device pointer of the internal framebuffer ring slot written by the last pm3d_render */
void *pm3d_last_frame_device(pm3d_ctx *ctx);

/* ------------------------------------------------------------------ presentation step
 * What GameScene.Draw does with rayScreen right after Camera.Draw (game/scene_game.go:169-185): scale it by
 * 1/renderScale (nearest) onto the screen-sized image and copy it, or run the CRT shader over it
 * (game/resources/shaders/crt.kage via game/render/effects/crt.go:40-53).  One fused pass, float32 like the
 * shader; HUD, transitions and menus stay with the game. */
enum { PM3D_PRESENT_COPY = 0, PM3D_PRESENT_CRT = 1, PM3D_PRESENT_CRT_CURVE = 2 };
/* screen size and g.renderScale (renderWidth = floor(screenWidth * renderScale), game/config.go:239-251);
 * call after pm3d_set_view */
int pm3d_set_present(pm3d_ctx *ctx, int out_w, int out_h, double render_scale);
/* d_src: n_views frames of the view size in device memory, or NULL for the frames of the last render.
 * out: n_views * out_h * out_w * 4 bytes (host, or device with PM3D_OUT_DEVICE). */
int pm3d_present(pm3d_ctx *ctx, const uint8_t *d_src, int n_views, int mode, uint8_t *out, uint32_t flags);

/* ------------------------------------------------------------------ LOS / collision path */

/* g.collisionMap = Map.GenerateWallCollisionLines(clipDistance) (game/mission.go:40):
 * n segments as x1,y1,x2,y2 */
int pm3d_upload_segments(pm3d_ctx *ctx, const double *xyxy, int n);
/* Game.lineOfSight (game/collision.go:40-65) for n rays (srcX,srcY,tgtX,tgtY):
 * out[i] = 1 iff no collision-map segment intersects ray i. */
int pm3d_los_batch(pm3d_ctx *ctx, const double *rays, int64_t n, uint8_t *out, uint32_t flags);
/* Wall part of Game.getValidMove (game/collision.go:69-97, 202-203, 307-342) for n moves
 * (posX,posY,newX,newY), checkAlternate=false, no sprite colliders: out_pos = x,y per move. */
int pm3d_move_batch(pm3d_ctx *ctx, const double *moves, int64_t n, double clip_distance, double *out_pos,
                    uint8_t *out_collided, uint32_t flags);

/* ---- full Game.getValidMove (game/collision.go:69-394) for a batch of moving entities.
 * The per-tick callers are game/combat.go:333 (every projectile, one goroutine each,
 * game/combat.go:206-239), game/update_sprites.go:263 and game/update_units.go:24. */

/* The moving model.Entity and the move it asks for (collision.go:69-81). */
typedef struct pm3d_mover {
    double pos_x, pos_y, pos_z;    /* entity.Pos(), entity.PosZ()                     */
    double move_x, move_y, move_z; /* moveX, moveY, moveZ                             */
    double radius, height;         /* CollisionRadius(), CollisionHeight()            */
    int32_t self, parent;          /* collider index of the entity itself / of entity.Parent(); -1 = none */
    uint8_t anchor, destroyed;     /* Anchor(), IsDestroyed()                         */
    uint8_t _pad[6];
} pm3d_mover;                      /* 80 bytes */

/* One collidable entity: g.player.Unit or a sprite of a collision type (collision.go:20-29). */
typedef struct pm3d_collider {
    double x, y, z, radius, height; /* Pos(), PosZ(), CollisionRadius(), CollisionHeight() */
    uint8_t anchor, destroyed;
    uint8_t _pad[6];
} pm3d_collider;                   /* 48 bytes */

/* What getValidMove returns, flattened: (*geom.Vector2, posZ, isCollision, collisions[0], len(collisions))
 * plus the closest intersection point the checkAlternate branch steers away from (collision.go:222-229). */
typedef struct pm3d_move_result {
    double x, y, z;
    int32_t collided;
    int32_t n_collisions;          /* len(collisionEntities)                          */
    int32_t first_entity;          /* collisions[0].entity as a collider index, -1 = none */
    int32_t n_points;              /* len(intersectPoints)                            */
    double first_x, first_y, first_z; /* collisions[0].collision, .collisionZ          */
    double near_x, near_y;         /* intersectPoints[minI]; NaN when there is none    */
} pm3d_move_result;                /* 80 bytes */

/* g.player.Unit and the collidable sprites, in the order the host walks them (the reference's
 * sync.Map order is unspecified; ties in the collision sort keep this order).  player_index = the
 * element that is g.player.Unit, -1 = none. */
int pm3d_upload_colliders(pm3d_ctx *ctx, const pm3d_collider *colliders, int n, int player_index);
/* getValidMove(entity, x, y, z, checkAlternate=false) for n movers against the uploaded collision
 * segments, colliders and level-0 grid.  The checkAlternate=true escape heuristics (collision.go:204-306,
 * trigonometry on the entity heading) are composed on the host from this call: see INTEGRATION.md. */
int pm3d_valid_move_batch(pm3d_ctx *ctx, const pm3d_mover *movers, int64_t n, double clip_distance,
                          pm3d_move_result *out, uint32_t flags);

/* Host mirror of Map.GenerateWallCollisionLines (game/model/map.go:811-851) so synthetic
 * workloads can build g.collisionMap without the Go side.  level0: [map_w][map_h].
 * Returns the number of segments (writes at most cap), or a negative status. */
int pm3d_generate_wall_collision_lines(const uint16_t *level0, int map_w, int map_h, int has_boundary_wall,
                                       double clip_distance, double *out_xyxy, int cap);

/* ------------------------------------------------------------------ clutter (ground cover around a camera)
 * ClutterHandler.Update (game/clutter.go:25-117; called per tick from game/scene_game.go:104-105 and once from
 * game/mission.go:101-103): the cells a fan of sample rays in front of the camera lands on (clutter.go:45-52) get
 * their rocks and shrubs from Go's math/rand seeded with Map.Seed + posId (clutter.go:75), so a cell's clutter is
 * the same whenever it comes into view.  The batch call evaluates that for n cameras at once; the sprites it
 * returns are what ClutterHandler.sprites holds after Update and what getRaycastSprites appends to the camera's
 * sprite list (game/sprites.go:121-124).  The handler's spritesByPosition cache is an optimisation of the
 * reference (a cell's sprites do not depend on the history) and stays a host-side detail. */

/* One model.MapClutter entry (game/model/map.go:128-133) as the host resolves it. */
typedef struct pm3d_clutter_def {
    double frequency;           /* Frequency: chance of this clutter in a matching cell */
    double scale;               /* Height / model.METERS_PER_UNIT (clutter.go:88)    */
    int32_t sheet;              /* g.tex.TextureImage(Image) as a sprite-sheet id    */
    int32_t rx0, ry0, rx1, ry1; /* GetSpriteSheetSlices(img, 1, 1): (0, 0, w-1, h-1) */
    int32_t _pad;
} pm3d_clutter_def;             /* 40 bytes */

/* One camera.  The sample angles are those of clutter.go:44-45 (heading -+ fov/2 in fov/20 steps: 21, or 20 when
 * the accumulated angle ends a rounding error past the limit); their cos/sin stay on the host like every
 * transcendental of this ABI. */
#define PM3D_CLUTTER_MAX_ANGLES 21
typedef struct pm3d_clutter_query {
    double cam_x, cam_y;        /* g.player.CameraPosition()                         */
    int32_t n_angles, _pad;
    double cs[PM3D_CLUTTER_MAX_ANGLES][2]; /* (cos a, sin a) per sample angle         */
} pm3d_clutter_query;           /* 360 bytes */

/* One clutter sprite: sprites.NewSprite(model.BasicVisualEntity(x, y, 0, AnchorBottom), scale, img)
 * (clutter.go:89-99). */
typedef struct pm3d_clutter_sprite {
    double x, y;
    int32_t cell_x, cell_y;     /* the cell it belongs to (posId = (cell_x-1)*mapWidth + cell_y) */
    int32_t def, _pad;          /* index into Map.Clutter                            */
} pm3d_clutter_sprite;          /* 32 bytes */

#define PM3D_CLUTTER_MAX_DEFS 64
#define PM3D_CLUTTER_MAX_DISTANCE 190.0
/* Mission-time clutter state (needs pm3d_upload_grid first).  defs: Map.Clutter, n_defs <= 64.  cell_class
 * [map_w][map_h]: a small id of the floor texture path TextureHandler.FloorTexturePathAt returns for the cell
 * (game/texture/texture.go:135-143); class_mask[c] bit i = defs[i].FloorPathMatch is nil or matches the path of
 * class c (the regular expressions run on the host, once per distinct path).  map_seed: Map.Seed
 * (game/model/map.go:93). */
int pm3d_upload_clutter(pm3d_ctx *ctx, const pm3d_clutter_def *defs, int n_defs, const uint8_t *cell_class,
                        const uint64_t *class_mask, int n_classes, int64_t map_seed);
/* ClutterHandler.Update for n cameras; clutter_distance = g.clutterDistance in cells (<= 190).
 *   out_count[n]: clutter sprites of each camera; out_records / out_sprites: [n][max_per_view], in creation order
 *   (sample angle, distance, clutter index); out_sprites (may be NULL) are the same sprites as ready-made
 *   pm3d_sprite records (z 0, anchor bottom, focusable, illumination 0) for pm3d_render.  A camera with more
 *   than max_per_view sprites reports its full count and keeps the first max_per_view.
 *   flags: PM3D_IN_DEVICE / PM3D_OUT_DEVICE (all three outputs) / PM3D_NO_SYNC. */
int pm3d_clutter_batch(pm3d_ctx *ctx, const pm3d_clutter_query *queries, int n, double clutter_distance,
                       int max_per_view, int32_t *out_count, pm3d_clutter_sprite *out_records,
                       pm3d_sprite *out_sprites, uint32_t flags);
/* The rngCooked table of Go's math/rand (607 words) as the library regenerates it (gen_cooked.go's 7.8e12 steps
 * by polynomial jump-ahead); exported so a test can pin it against known answers. */
int pm3d_go_rand_cooked(uint64_t *out607);

/* ------------------------------------------------------------------ per-frame sprite list
 * Game.getRaycastSprites (game/sprites.go:102-145): which sprites a camera's Update gets.  For every camera:
 * the map's sprites that pass `renderDistance < 0 || sprite is the player's target || PointInProximity(
 * renderDistance, cam, sprite)` (:113-114; the Chebyshev test of game/model/math.go:93-96 with ceil(distance)), in
 * the order the host walks g.sprites, then all of its clutter (:121-124), then the extras the game appends
 * (reticle lead, current nav point, the player sprite under a debug camera, :127-142).  The output is exactly the
 * `sprites` + `sprite_offset` pair pm3d_render / pm3d_render_resident take, so a headless batch can go from camera
 * poses to frames without the sprite lists ever leaving HBM. */
typedef struct pm3d_list_query {
    double cam_x, cam_y;        /* g.player.CameraPosXY()                             */
    int32_t target;             /* index of g.player.Target() in `world`, -1 = none   */
    int32_t _pad;
} pm3d_list_query;              /* 24 bytes */
/*   world[n_world]: the map's sprites, shared by all cameras.
 *   clutter (may be NULL): [n][clutter_stride] records with clutter_count[n] valid ones per camera — the out_sprites /
 *   out_count of pm3d_clutter_batch.   extras (may be NULL): per-camera extra sprites concatenated, extra_offset[n+1].
 *   out_sprites[max_total], out_offset[n+1]; out_index[max_total] (may be NULL): where each listed sprite came from —
 *   world index, -1-k for clutter sprite k of that camera, INT32_MIN+k for its extra k (the host maps the rectangles
 *   pm3d_render returns back to its Sprite objects with it).  *out_total (host, may be NULL) = out_offset[n].  More
 *   than max_total sprites is PM3D_ERR_CAPACITY (nothing is written past max_total).
 *   flags: PM3D_IN_DEVICE (world, queries, clutter*, extras* are device memory), PM3D_OUT_DEVICE (out_sprites,
 *   out_offset, out_index), PM3D_NO_SYNC (enqueue only: needs PM3D_OUT_DEVICE; out_total is not written, the caller
 *   reads out_offset[n] when it needs the total; more than max_total sprites cuts the lists at max_total and the
 *   next pm3d_sync returns PM3D_ERR_CAPACITY). */
int pm3d_raycast_sprites_batch(pm3d_ctx *ctx, const pm3d_sprite *world, int n_world, const pm3d_list_query *queries,
                               int n, double render_distance, const pm3d_sprite *clutter, const int32_t *clutter_count,
                               int clutter_stride, const pm3d_sprite *extras, const int32_t *extra_offset,
                               int max_total, pm3d_sprite *out_sprites, int32_t *out_offset, int32_t *out_index,
                               int32_t *out_total, uint32_t flags);

/* ------------------------------------------------------------------ multi-GPU: sharded batches + gathers
 * the path shards without an exchange step.  Unit cameras are independent (one per AI unit,
 * game/ai_combat.go / game/update_units.go drive one entity each), so a batch of views is dealt round-robin
 * (view i -> rank i % n); LOS rays are independent (game/collision.go:40-65), so a batch is cut into contiguous
 * ranges; grid, atlases, sheets and collision segments are replicated on every GPU (a few MB: upload them to every
 * ctx).  What travels is the COMPACT result of every unit — convergence record + 64-bit frame checksum per view,
 * one byte per ray — gathered to EVERY rank over NVLink by a collective the library enqueues on the ctx stream right
 * behind the kernels (no host round trip between compute and gather).  The framebuffers stay where they were
 * rendered.  The library owns the NCCL communicator (libnccl.so.2 is dlopen'ed on first use, PM3D_NCCL_LIB
 * overrides the name), so the Go host needs no NCCL binding:
 *   one process per GPU:  rank 0 calls pm3d_comm_unique_id and hands the 128 bytes to the other ranks (any channel
 *                         the host has), every rank calls pm3d_comm_init on its ctx;
 *   one process, n GPUs:  pm3d_group_create makes one ctx per device plus the communicator over them. */
#define PM3D_COMM_ID_BYTES 128
#define PM3D_GROUP_MAX 8
int pm3d_comm_unique_id(uint8_t *id128);
int pm3d_comm_init(pm3d_ctx *ctx, const uint8_t *id128, int rank, int n_ranks); /* collective: every rank calls it */
int pm3d_comm_destroy(pm3d_ctx *ctx);                                           /* also done by pm3d_destroy     */
int pm3d_comm_rank(const pm3d_ctx *ctx);
int pm3d_comm_size(const pm3d_ctx *ctx);                                        /* 1 without a communicator      */
/* the sharding arithmetic: how many of n_views rank owns (views rank, rank + n, ...); the ray range [lo, hi) */
int64_t pm3d_shard_view_count(int64_t n_views, int rank, int n_ranks);
int pm3d_shard_ray_range(int64_t n_rays, int rank, int n_ranks, int64_t *lo, int64_t *hi);
/* raw all-gather of bytes_per_rank bytes from every rank (device memory, rank-major in d_recv), enqueue only */
int pm3d_comm_all_gather(pm3d_ctx *ctx, const void *d_send, void *d_recv, int64_t bytes_per_rank);
/* pm3d_render_resident for THIS rank's n_local views of a batch of n_views_global (its views in shard order:
 * global view rank + k * n_ranks is local view k), then the gather: on return (enqueue only, pm3d_sync to wait)
 * d_all_conv[n_views_global] and d_all_checksums[n_views_global] (device memory, either may be NULL) hold the
 * convergence record and the frame checksum of EVERY view of the batch, in global view order, on every rank.
 * The checksum / gather tail runs on the ctx stream behind the batch; a FOLLOWING frame batch does not wait for it
 * (nothing the tail writes is a frame input), so its cast / setup launch overlaps this batch's compose and gather. */
int pm3d_comm_render_views(pm3d_ctx *ctx, const pm3d_view *d_views, int n_local, const pm3d_sprite *d_sprites,
                           const int32_t *d_sprite_offset, int n_sprites_total, uint8_t *d_out_rgba, int n_views_global,
                           pm3d_convergence *d_all_conv, uint64_t *d_all_checksums);
/* pm3d_los_batch for THIS rank's contiguous range of a batch of n_global rays (d_rays holds rays [lo, hi) of
 * pm3d_shard_ray_range), written straight into d_all_out + lo, then the gather: d_all_out[n_global] holds every
 * ray's answer on every rank (enqueue only).  The answers cross the links as bits (1/8 of the bytes) when n_global is a
 * multiple of 32 * n_ranks and d_all_out is 16-byte aligned, as bytes otherwise: same result either way. */
int pm3d_comm_los_batch(pm3d_ctx *ctx, const double *d_rays, int64_t n_local, int64_t n_global, uint8_t *d_all_out);

/* One process driving n GPUs (what a single Go game-server process does). */
typedef struct pm3d_group pm3d_group;
int pm3d_group_create(const int *devices, int n, pm3d_group **out);
void pm3d_group_destroy(pm3d_group *g);
const char *pm3d_group_last_error(const pm3d_group *g); /* g may be NULL for pm3d_group_create failures */
int pm3d_group_size(const pm3d_group *g);
pm3d_ctx *pm3d_group_ctx(pm3d_group *g, int i);         /* for the replicated uploads, pm3d_set_view, ... */
int pm3d_group_sync(pm3d_group *g);
/* HOST arrays in, compact results of all views out (host): the views are dealt i % n, rendered on their device
 * (frames stay in that device's frame buffer), gathered, and read back from device 0.  Synchronous. */
int pm3d_group_render_views(pm3d_group *g, const pm3d_view *views, int n_views, const pm3d_sprite *sprites,
                            const int32_t *sprite_offset, pm3d_convergence *out_conv, uint64_t *out_checksums);
/* Game.lineOfSight for n host rays cut into contiguous ranges over the group's devices.  Synchronous. */
int pm3d_group_los_batch(pm3d_group *g, const double *rays, int64_t n_rays, uint8_t *out);

/* ------------------------------------------------------------------ device memory + timing
 * (so a cgo / ctypes host needs no CUDA binding of its own) */
int pm3d_device_alloc(pm3d_ctx *ctx, int64_t bytes, void **out);
int pm3d_device_free(pm3d_ctx *ctx, void *ptr);
int pm3d_host_alloc(pm3d_ctx *ctx, int64_t bytes, void **out); /* pinned */
int pm3d_host_free(pm3d_ctx *ctx, void *ptr);
int pm3d_memcpy_h2d(pm3d_ctx *ctx, void *dst, const void *src, int64_t bytes);
int pm3d_memcpy_d2h(pm3d_ctx *ctx, void *dst, const void *src, int64_t bytes);
/* CUDA-event timers on the ctx stream: slot in [0, 16) */
int pm3d_timer_start(pm3d_ctx *ctx, int slot);
int pm3d_timer_stop(pm3d_ctx *ctx, int slot);
int pm3d_timer_elapsed_ms(pm3d_ctx *ctx, int slot, float *ms); /* syncs on the stop event */
/* Per-kernel CUDA-event timing of the frame path: after pm3d_ktimer_begin(ctx, max_frames) every
 * render records an event pair around each of its kernels (0 = front: column cast + view
 * setup, 1 = unused (reads 0), 2 = compose) until max_frames frames were sampled.  pm3d_ktimer_collect syncs, writes the
 * per-frame durations in ms as out_ms[frame*3 + kernel], returns the number of frames sampled
 * and stops the sampling. */
int pm3d_ktimer_begin(pm3d_ctx *ctx, int max_frames);
int pm3d_ktimer_collect(pm3d_ctx *ctx, float *out_ms, int cap_frames);
/* kernels launched by this ctx since creation (bench.py "gpu_launches") */
int64_t pm3d_launch_count(const pm3d_ctx *ctx);
/* write `bytes` of a scratch buffer (> L2) to flush the L2 between timed iterations */
int pm3d_flush_l2(pm3d_ctx *ctx);
/* Device self-check of the division-free quotient the floor cast uses for weight = currentDist / distWall
 * (Go: one float64 division per floor pixel): blocks*256 threads each compare `iters` operand pairs
 * (random, adversarial and path-range) bit for bit against the IEEE division.  out_mismatches must read 0. */
int pm3d_selfcheck_division(pm3d_ctx *ctx, uint64_t seed, int blocks, int iters, uint64_t *out_mismatches,
                            uint64_t *out_tested);
/* Device self-check of the fast floor cast (texel index from one fused multiply-add plus a rounding-boundary
 * test) against the reference's float64 sequence int((weight*floorXWall + (1-weight)*posX) * texSize), on random
 * operands and on operands steered next to integer texel boundaries.  out_accepted counts the pairs the fast
 * path accepted (the rest fall back to the reference sequence in the kernel); out_mismatches must read 0. */
int pm3d_selfcheck_floor(pm3d_ctx *ctx, uint64_t seed, int blocks, int iters, uint64_t *out_mismatches,
                         uint64_t *out_tested, uint64_t *out_accepted);

/* ---- Pathing.FindPath (game/model/pathing.go:123-186; call site game/ai_movement.go:64), batched: one query per AI
 * unit.  queries[n][4] = start x, y, finish x, y (float64 map positions; the tiles are int(x), int(y) as in the
 * reference).  The search runs over level 0 of the grid given to pm3d_upload_grid (initPathing, pathing.go:94-112: a
 * tile is a blocker iff its cell is non-zero) and reproduces astar.Path of github.com/harbdog/go-astar — including the
 * container/heap order that picks one of the many equally short paths — then the reference's reversal into cell
 * centres and its Bezier smoothing in segments of seven steps (game/common/bezier).
 * out_points[n][max_points][2]: the points FindPath returns, start -> finish.  out_count[n]: how many there are
 * (more than max_points: only the first max_points were written), -1 = "unable to find path" (the reference's error;
 * also a start or finish outside the map).  flags: PM3D_IN_DEVICE / PM3D_OUT_DEVICE / PM3D_NO_SYNC (device buffers
 * only).  Scratch: 16 bytes per map cell and search in flight (as many searches at once as a third of the free
 * device memory holds, at most 48 GiB; more queries run in waves), owned by the ctx until pm3d_destroy. */
int pm3d_find_path_batch(pm3d_ctx *ctx, const double *queries, int64_t n, int max_points, double *out_points,
                         int32_t *out_count, uint32_t flags);

/* Device self-check of the wall cast's closed-form DDA jump (pm3d_internal.cuh dda_jump): blocks*256 rays (random
 * poses and directions, exact diagonals, rational slopes, cell-centre / lattice positions) are walked `steps` steps
 * by the reference's sequential loop; at random step counts the certified closed-form state (x steps, y steps, side
 * of the last step) must equal the loop's.  out_certified counts the checks the certificate accepted (the rest
 * take the sequential loop in the kernel); out_mismatches must read 0. */
int pm3d_selfcheck_dda(pm3d_ctx *ctx, uint64_t seed, int blocks, int steps, uint64_t *out_mismatches,
                       uint64_t *out_tested, uint64_t *out_certified);

#ifdef __cplusplus
}
#endif
#endif
