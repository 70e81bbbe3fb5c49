// camera_demo.cpp — drives libpm3d.so through the C++ mirror of raycaster.Camera (raycaster_camera.hpp)
// exactly like game/mission.go:74-97 + game/scene_game.go:161-167 do: NewCamera, setters, Update, Draw.
// Writes the frame (h*w*4 premultiplied RGBA8) to argv[1]; tests/test_host_cpp_gpu.py re-creates the same
// scene through the oracle and compares every byte.
#include <cstdio>
#include <cstring>

#include "raycaster_camera.hpp"

using namespace raycaster;

static Image make_tile(int size, int seed, bool holes) {
    Image im;
    im.W = im.H = size;
    im.Pix.resize((size_t)4 * size * size);
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++) {
            uint8_t *p = &im.Pix[((size_t)y * size + x) * 4];
            p[0] = (uint8_t)((x * 3 + seed * 40) & 255);
            p[1] = (uint8_t)((y * 2 + seed * 70) & 255);
            p[2] = (uint8_t)(((x ^ y) + seed * 20) & 255);
            p[3] = 255;
            if (holes && (x % 32) > 20 && (y % 32) > 20) p[0] = p[1] = p[2] = p[3] = 0;
        }
    return im;
}

struct DemoMap : Map {
    std::vector<std::vector<std::vector<int>>> levels;
    int NumLevels() const override { return 2; }
    const std::vector<std::vector<int>> &Level(int l) const override { return levels[l < (int)levels.size() ? l : (int)levels.size() - 1]; }
};

struct DemoTex : TextureHandler {
    const DemoMap *m;
    Image wall1, wall2, wall2x, floorA, floorB;
    const Image *TextureAt(int x, int y, int l, int side) const override {
        const auto &lv = m->Level(l);
        if (x < 0 || y < 0 || x >= (int)lv.size() || y >= (int)lv[0].size()) return nullptr;
        const int v = lv[x][y];
        if (v <= 0) return nullptr;
        if (v == 2) return side == 1 ? &wall2x : &wall2;  // SideX image on side 1 (map.go:110)
        return &wall1;
    }
    const Image *FloorTextureAt(int x, int y) const override {
        if (x < 0 || y < 0) return nullptr;
        return (x + y) % 3 == 0 ? &floorB : &floorA;
    }
};

struct DemoSprite : Sprite {
    Vector2 pos; double z = 0, scale = 1; const Image *img; Rectangle rect; bool has_rect = false; Rectangle screen;
    Vector2 Pos() const override { return pos; }
    double PosZ() const override { return z; }
    double Scale() const override { return scale; }
    SpriteAnchor VerticalAnchor() const override { return AnchorBottom; }
    const Image *Texture() const override { return img; }
    Rectangle TextureRect() const override { return rect; }
    double Illumination() const override { return 10.0; }
    void SetScreenRect(const Rectangle *r) override { has_rect = r != nullptr; if (r) screen = *r; }
    bool IsFocusable() const override { return true; }
};

int main(int argc, char **argv) {
    const int W = 320, H = 200, T = 64;
    DemoMap m;
    m.levels.assign(2, std::vector<std::vector<int>>(12, std::vector<int>(10, 0)));
    for (int x = 0; x < 12; x++) { m.levels[0][x][0] = 1; m.levels[0][x][9] = 1; }
    for (int y = 0; y < 10; y++) { m.levels[0][0][y] = 1; m.levels[0][11][y] = 1; }
    m.levels[0][7][4] = 2; m.levels[1][7][4] = 2; m.levels[0][8][6] = 1;
    DemoTex tex;
    tex.m = &m;
    tex.wall1 = make_tile(T, 1, false); tex.wall2 = make_tile(T, 2, true); tex.wall2x = make_tile(T, 3, false);
    tex.floorA = make_tile(T, 4, false); tex.floorB = make_tile(T, 5, false);
    Image sky = make_tile(T, 6, false), box = make_tile(T, 7, false);
    Image sheet = make_tile(48, 8, true);
    sheet.H = 40; sheet.Pix.resize((size_t)4 * 48 * 40);

    try {
        Camera cam(W, H, T, m, tex);                       // game/mission.go:74
        cam.SetRenderDistance(50.0);                       // :75
        cam.SetAlwaysSetSpriteScreenRect(true);            // :76
        cam.SetFloorTexture(&box); cam.SetSkyTexture(&sky);  // :79, :82
        cam.SetLightFalloff(-60.0); cam.SetGlobalIllumination(120.0);  // :90-91
        cam.SetLightRGB(RGB{40, 50, 60}, RGB{255, 240, 230});          // :92
        cam.SetPosition(Vector2{3.3, 4.6}); cam.SetPositionZ(0.6);     // game/camera.go:31-32
        cam.SetHeadingAngle(0.2); cam.SetPitchAngle(0.05);             // :33-34
        cam.SetFovAngle(70.0, 1.0);                                    // game/config.go:290
        DemoSprite a, b;
        a.pos = {5.5, 5.0}; a.scale = 0.7; a.img = &sheet; a.rect = {0, 0, 48, 40};
        b.pos = {6.2, 3.9}; b.scale = 0.5; b.z = 0.3; b.img = &sheet; b.rect = {4, 2, 40, 38};
        std::vector<Sprite *> sprites{&a, &b};
        cam.Update(sprites);                               // game/scene_game.go:161
        Sprite *conv = cam.GetConvergenceSprite();         // :164
        const std::vector<uint8_t> &frame = cam.Draw();    // :167
        if (argc > 1) {
            FILE *f = fopen(argv[1], "wb");
            if (!f) { perror("fopen"); return 2; }
            fwrite(frame.data(), 1, frame.size(), f);
            fclose(f);
        }
        printf("frame %dx%d, convergence sprite %s, rect a %s [%d %d %d %d], fov_v %.6f\n", W, H,
               conv == &a ? "a" : (conv == &b ? "b" : "none"), a.has_rect ? "set" : "nil", a.screen.MinX, a.screen.MinY,
               a.screen.MaxX, a.screen.MaxY, cam.FovRadiansVertical());
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
