// raycaster_camera.hpp — header-only C++ mirror of the raycaster.Camera surface the game drives
// (SURVEY.md §8a K1-K5, §8b), on top of the C-ABI of include/pm3d.h.
//
// The reference's host language is Go; this image has no Go toolchain, so the host side above the
// C-ABI is mirrored in C++ (the reference is compiled code) with the Go names and argument meaning:
//   raycaster.NewCamera(w, h, texSize, Map, TextureHandler)   game/mission.go:74
//   Camera.Update(sprites) / GetConvergenceSprite() / Draw()   game/scene_game.go:161-167
//   setters / getters                                          game/camera.go:31-52, game/config.go:248-291,
//                                                              game/input.go:764-792, game/mission.go:75-92
// The Map / TextureHandler / Sprite callbacks are abstract classes with the reference's method sets
// (game/model/map.go:253-264, game/texture/texture.go:69-133, game/render/sprites/sprite.go:49-91); they are
// resolved BEFORE the device call: once per mission for grid and textures, once per frame for the sprites.
// Camera methods cannot fail at the reference's call sites, so errors throw (the Go shim would log.Fatal).
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "pm3d.h"

namespace raycaster {

enum SpriteAnchor { AnchorBottom = PM3D_ANCHOR_BOTTOM, AnchorCenter = PM3D_ANCHOR_CENTER, AnchorTop = PM3D_ANCHOR_TOP };

struct Vector2 { double X = 0, Y = 0; };
struct Rectangle { int MinX = 0, MinY = 0, MaxX = 0, MaxY = 0; };
struct RGB { uint8_t R = 0, G = 0, B = 0; };

// premultiplied RGBA8 image, row-major (what *ebiten.Image / *image.RGBA hold)
struct Image { int W = 0, H = 0; std::vector<uint8_t> Pix; };

// raycaster.Map (game/model/map.go:253-264)
class Map {
public:
    virtual ~Map() = default;
    virtual int NumLevels() const = 0;
    // level grid in [x][y] order; levels above the stored ones repeat the top one
    virtual const std::vector<std::vector<int>> &Level(int levelNum) const = 0;
};

// raycaster.TextureHandler (game/texture/texture.go:69-97, 119-133)
class TextureHandler {
public:
    virtual ~TextureHandler() = default;
    virtual const Image *TextureAt(int x, int y, int levelNum, int side) const = 0;
    virtual const Image *FloorTextureAt(int x, int y) const = 0;
};

// raycaster.Sprite (game/render/sprites/sprite.go:49-91)
class Sprite {
public:
    virtual ~Sprite() = default;
    virtual Vector2 Pos() const = 0;
    virtual double PosZ() const = 0;
    virtual double Scale() const = 0;
    virtual SpriteAnchor VerticalAnchor() const = 0;
    virtual const Image *Texture() const = 0;   // the sheet; TextureRect() selects the cell
    virtual Rectangle TextureRect() const = 0;
    virtual double Illumination() const = 0;
    virtual void SetScreenRect(const Rectangle *rect) = 0;  // nullptr = not on screen
    virtual bool IsFocusable() const = 0;
};

class Camera {
public:
    // raycaster.NewCamera: flattens the callbacks and uploads the mission data
    Camera(int w, int h, int texSize, const Map &m, const TextureHandler &tex, int device = 0)
        : w_(w), h_(h), texSize_(texSize) {
        check(nullptr, pm3d_create(device, nullptr, &ctx_));
        const int levels = m.NumLevels();
        const auto &l0 = m.Level(0);
        const int mw = (int)l0.size(), mh = mw ? (int)l0[0].size() : 0;
        std::vector<uint16_t> grid;
        grid.reserve((size_t)levels * mw * mh);
        std::vector<int32_t> idx, s0, s1;
        std::map<int, bool> seen;
        for (int l = 0; l < levels; l++) {
            const auto &lv = m.Level(l);
            for (int x = 0; x < mw; x++)
                for (int y = 0; y < mh; y++) {
                    const int v = lv[x][y];
                    grid.push_back((uint16_t)v);
                    if (v > 0 && !seen[v]) {
                        seen[v] = true;
                        idx.push_back(v);
                        s0.push_back(tile_of(tex.TextureAt(x, y, l, 0)));
                        s1.push_back(tile_of(tex.TextureAt(x, y, l, 1)));
                    }
                }
        }
        std::vector<int16_t> floor((size_t)mw * mh);
        for (int x = 0; x < mw; x++)
            for (int y = 0; y < mh; y++) floor[(size_t)x * mh + y] = (int16_t)tile_of(tex.FloorTextureAt(x, y));
        if (atlas_.empty()) atlas_.assign((size_t)4 * texSize * texSize, 0);
        check(ctx_, pm3d_upload_grid(ctx_, grid.data(), levels, mw, mh, levels));
        check(ctx_, pm3d_upload_wall_textures(ctx_, idx.data(), s0.data(), s1.data(), (int)idx.size()));
        check(ctx_, pm3d_upload_floor_tiles(ctx_, floor.data(), mw, mh));
        upload_atlas();
        SetViewSize(w, h);
    }
    ~Camera() { pm3d_destroy(ctx_); }
    Camera(const Camera &) = delete;
    Camera &operator=(const Camera &) = delete;

    // ---- K5 (game/camera.go:31-34, game/config.go:249-290, game/input.go:764-792, game/mission.go:75-92)
    void SetPosition(const Vector2 &p) { pos_ = p; }
    void SetPositionZ(double z) { posZ_ = z; }
    void SetHeadingAngle(double a) { heading_ = a; }
    void SetPitchAngle(double a) { pitch_ = a; }
    void SetFovAngle(double fovDegrees, double fovDepth) { fov_ = fovDegrees * M_PI / 180.0; fovDepth_ = fovDepth; }
    void SetRenderDistance(double d) { renderDistance_ = d; }
    void SetLightFalloff(double f) { falloff_ = f; }
    void SetGlobalIllumination(double g) { illumination_ = g; }
    void SetLightRGB(RGB mn, RGB mx) { minRGB_ = mn; maxRGB_ = mx; }
    void SetAlwaysSetSpriteScreenRect(bool) {}
    void SetFloorTexture(const Image *img) { floorBox_ = tile_of(img); upload_atlas(); check(ctx_, pm3d_set_box_tiles(ctx_, sky_, floorBox_)); }
    void SetSkyTexture(const Image *img) { sky_ = tile_of(img); upload_atlas(); check(ctx_, pm3d_set_box_tiles(ctx_, sky_, floorBox_)); }
    void SetViewSize(int w, int h) {
        w_ = w; h_ = h;
        check(ctx_, pm3d_set_view(ctx_, w, h, 1, 2048));
        frame_.assign((size_t)4 * w * h, 0);
    }
    Vector2 GetPosition() const { return pos_; }
    double GetPositionZ() const { return posZ_; }
    double FovDepth() const { return fovDepth_; }
    double FovRadians() const { return fov_; }
    double FovRadiansVertical() const { return 2.0 * std::atan(std::tan(fov_ / 2.0) * ((double)h_ / (double)w_)); }

    // ---- K2/K3/K4 (game/scene_game.go:161-167)
    void Update(const std::vector<Sprite *> &sprites) {
        sprites_ = sprites;
        std::vector<pm3d_sprite> snap(sprites.size());
        for (size_t i = 0; i < sprites.size(); i++) {
            const Sprite &s = *sprites[i];
            const Rectangle r = s.TextureRect();
            pm3d_sprite &o = snap[i];
            o = pm3d_sprite{};
            o.x = s.Pos().X; o.y = s.Pos().Y; o.z = s.PosZ();
            o.scale = s.Scale(); o.illumination = s.Illumination();
            o.sheet = sheet_of(s.Texture());
            o.rx0 = r.MinX; o.ry0 = r.MinY; o.rx1 = r.MaxX; o.ry1 = r.MaxY;
            o.anchor = (uint8_t)s.VerticalAnchor();
            o.focusable = s.IsFocusable() ? 1 : 0;
        }
        rects_.resize(sprites.size() + 1);
        const pm3d_view v = view();
        const int32_t off[2] = {0, (int32_t)sprites.size()};
        check(ctx_, pm3d_render(ctx_, &v, 1, snap.empty() ? nullptr : snap.data(), off, frame_.data(), rects_.data(), &conv_, 0));
        for (size_t i = 0; i < sprites.size(); i++) {
            const pm3d_rect &r = rects_[i];
            if (r.x0 > r.x1) sprites[i]->SetScreenRect(nullptr);
            else { Rectangle rc{r.x0, r.y0, r.x1, r.y1}; sprites[i]->SetScreenRect(&rc); }
        }
    }
    Sprite *GetConvergenceSprite() const { return conv_.sprite < 0 ? nullptr : sprites_[(size_t)conv_.sprite]; }
    double GetConvergenceDistance() const { return conv_.distance; }
    // Draw: the frame is already composed; returns h*w*4 premultiplied RGBA8
    const std::vector<uint8_t> &Draw() const { return frame_; }

    // ---- LOS / collision (game/collision.go:40-65, :69-97)
    void UploadSegments(const std::vector<double> &xyxy) { check(ctx_, pm3d_upload_segments(ctx_, xyxy.data(), (int)(xyxy.size() / 4))); }
    std::vector<uint8_t> LineOfSightBatch(const std::vector<double> &rays) {
        std::vector<uint8_t> out(rays.size() / 4);
        check(ctx_, pm3d_los_batch(ctx_, rays.data(), (int64_t)out.size(), out.data(), 0));
        return out;
    }

    // lowering of the camera state to what the device consumes (transcendentals stay on the host)
    pm3d_view view() const {
        pm3d_view v{};
        const double dx = fovDepth_ * std::cos(heading_), dy = fovDepth_ * std::sin(heading_), t = std::tan(fov_ / 2.0);
        long pitch = (long)((double)h_ * fovDepth_ * std::tan(pitch_));
        const long lo = -(h_ / 2), hi = (long)((double)h_ * fovDepth_);
        pitch = pitch < lo ? lo : (pitch > hi ? hi : pitch);
        v.pos_x = pos_.X; v.pos_y = pos_.Y; v.cam_z_px = (posZ_ - 0.5) * (double)h_;
        v.dir_x = dx; v.dir_y = dy; v.plane_x = dy * t; v.plane_y = -dx * t;
        v.pitch_px = (int32_t)pitch;
        v.render_distance = renderDistance_; v.light_falloff = falloff_; v.global_illumination = illumination_;
        v.min_light[0] = minRGB_.R; v.min_light[1] = minRGB_.G; v.min_light[2] = minRGB_.B; v.min_light[3] = 255;
        v.max_light[0] = maxRGB_.R; v.max_light[1] = maxRGB_.G; v.max_light[2] = maxRGB_.B; v.max_light[3] = 255;
        return v;
    }

private:
    static void check(pm3d_ctx *ctx, int rc) {
        if (rc != 0) throw std::runtime_error(std::string("pm3d: ") + pm3d_last_error(ctx));
    }
    int tile_of(const Image *img) {
        if (!img) return -1;
        auto it = tiles_.find(img);
        if (it != tiles_.end()) return it->second;
        if (img->W != texSize_ || img->H != texSize_) throw std::runtime_error("pm3d: texture must be texSize x texSize");
        const int t = (int)tiles_.size();
        tiles_[img] = t;
        atlas_.insert(atlas_.end(), img->Pix.begin(), img->Pix.end());
        return t;
    }
    void upload_atlas() {
        const size_t tile_bytes = (size_t)4 * texSize_ * texSize_;
        const int n = (int)(atlas_.size() / tile_bytes);
        check(ctx_, pm3d_upload_atlas(ctx_, atlas_.data(), n < 1 ? 1 : n, texSize_));
    }
    int sheet_of(const Image *img) {
        if (!img) return -1;
        auto it = sheets_.find(img);
        if (it != sheets_.end()) return it->second;
        const int id = (int)sheets_.size();
        sheets_[img] = id;
        sheet_off_.push_back((int64_t)sheet_px_.size());
        sheet_w_.push_back(img->W);
        sheet_h_.push_back(img->H);
        sheet_px_.insert(sheet_px_.end(), img->Pix.begin(), img->Pix.end());
        check(ctx_, pm3d_upload_sprite_sheets(ctx_, sheet_px_.data(), sheet_off_.data(), sheet_w_.data(), sheet_h_.data(), id + 1));
        return id;
    }

    pm3d_ctx *ctx_ = nullptr;
    int w_, h_, texSize_;
    Vector2 pos_{1, 1};
    double posZ_ = 0.5, heading_ = 0, pitch_ = 0, fov_ = 60.0 * M_PI / 180.0, fovDepth_ = 1.0;
    double renderDistance_ = -1, falloff_ = -100, illumination_ = 300;
    RGB minRGB_{0, 0, 0}, maxRGB_{255, 255, 255};
    int sky_ = -1, floorBox_ = -1;
    std::map<const Image *, int> tiles_, sheets_;
    std::vector<uint8_t> atlas_, sheet_px_, frame_;
    std::vector<int64_t> sheet_off_;
    std::vector<int32_t> sheet_w_, sheet_h_;
    std::vector<Sprite *> sprites_;
    std::vector<pm3d_rect> rects_;
    pm3d_convergence conv_{};
};

}  // namespace raycaster
