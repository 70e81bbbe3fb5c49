// go_harness — headless golden-vector dump of the reference's raycast for the C1 pose (UNVERIFIED: written without
// a Go toolchain; see README.md in this directory).
//
// It builds raycaster.NewCamera over the reference's own model.Map and texture.TextureHandler exactly as
// game/mission.go:74-97 does, wraps the TextureHandler in a recorder, runs Camera.Update for one pose and writes what
// is observable WITHOUT patching github.com/harbdog/raycaster-go:
//
//   * every TextureAt(x, y, level, side) call Camera.Update makes = the DDA hit cell and wall side of a column of a
//     level (the per-column slice records themselves are unexported fields of the library's Camera);
//   * every FloorTextureAt(x, y) call = the map cell of a floor pixel, in the library's own pixel order;
//   * the rectangle each sprite receives through SetScreenRect;
//   * GetConvergenceDistance / GetConvergencePoint / GetConvergenceSprite.
//
// Output: golden_c1.bin, little endian —
//   "PM3DGOLD" | int32 w, h, nLevels, nTex, nFloor, nSprites |
//   nTex   x (int32 x, y, level, side)                            in call order per level (level-major after sorting)
//   nFloor x (int32 x, y)                                         in call order
//   nSprites x (int32 x0, y0, x1, y1; -1 rect = nil)              in the order of the sprite list
//   float64 convergenceDistance, pointX, pointY; int32 convergenceSpriteIndex (-1 = nil)
// tests/golden/make_go_golden.py (this repo) turns it into an .npz next to the existing fixtures; once it exists,
// tests/test_oracle_cpu.py::test_go_golden_c1 compares the oracle's hit cells / sides / floor cells / rectangles /
// convergence with it and the "parity unpinned" notes can go.
package main

import (
	"encoding/binary"
	"image"
	"log"
	"math"
	"os"
	"sort"
	"sync"

	"github.com/hajimehoshi/ebiten/v2"
	"github.com/harbdog/raycaster-go"
	"github.com/harbdog/raycaster-go/geom"

	"github.com/pixelmek-3d/pixelmek-3d/game/model"
	"github.com/pixelmek-3d/pixelmek-3d/game/resources"
	"github.com/pixelmek-3d/pixelmek-3d/game/texture"
)

type texCall struct{ x, y, level, side int32 }
type floorCall struct{ x, y int32 }

// recorder forwards to the game's TextureHandler and remembers every call
type recorder struct {
	inner *texture.TextureHandler
	mu    sync.Mutex
	tex   []texCall
	floor []floorCall
}

func (r *recorder) TextureAt(x, y, levelNum, side int) *ebiten.Image {
	r.mu.Lock()
	r.tex = append(r.tex, texCall{int32(x), int32(y), int32(levelNum), int32(side)})
	r.mu.Unlock()
	return r.inner.TextureAt(x, y, levelNum, side)
}

func (r *recorder) FloorTextureAt(x, y int) *image.RGBA {
	r.mu.Lock()
	r.floor = append(r.floor, floorCall{int32(x), int32(y)})
	r.mu.Unlock()
	return r.inner.FloorTextureAt(x, y)
}

// goldSprite is the smallest raycaster.Sprite: a map sprite of the bundled map with its image's full rectangle
type goldSprite struct {
	pos    geom.Vector2
	z      float64
	scale  float64
	anchor raycaster.SpriteAnchor
	img    *ebiten.Image
	rect   *image.Rectangle
}

func (s *goldSprite) Pos() *geom.Vector2                     { return &s.pos }
func (s *goldSprite) PosZ() float64                          { return s.z }
func (s *goldSprite) Scale() float64                         { return s.scale }
func (s *goldSprite) VerticalAnchor() raycaster.SpriteAnchor { return s.anchor }
func (s *goldSprite) Texture() *ebiten.Image                 { return s.img }
func (s *goldSprite) TextureRect() image.Rectangle           { return s.img.Bounds() }
func (s *goldSprite) Illumination() float64                  { return 0 }
func (s *goldSprite) SetScreenRect(r *image.Rectangle)       { s.rect = r }
func (s *goldSprite) IsFocusable() bool                      { return true }

func main() {
	const w, h = 640, 360 // BASELINE config C1
	mission, err := model.LoadMission("trial_day.yaml")
	if err != nil {
		log.Fatal(err)
	}
	m := mission.Map()
	tex := texture.NewTextureHandler(m)
	tex.SetRenderFloorTex(true)
	rec := &recorder{inner: tex}

	camera := raycaster.NewCamera(w, h, resources.TexSize, m, rec) // game/mission.go:74
	camera.SetRenderDistance(100)                                   // 2000 m / 20 m per cell (game/config.go:142)
	camera.SetAlwaysSetSpriteScreenRect(true)
	camera.SetLightFalloff(m.Lighting.Falloff)
	camera.SetGlobalIllumination(m.Lighting.Illumination)
	minL, maxL := m.Lighting.LightRGB()
	camera.SetLightRGB(*minL, *maxL)

	// trial_day drop zone, cockpit height 0.5, pitch 0, FOV 70 (game/config.go:287-292)
	pos := &geom.Vector2{X: mission.DropZone.Position[0], Y: mission.DropZone.Position[1]}
	camera.SetPosition(pos)
	camera.SetPositionZ(0.5)
	camera.SetHeadingAngle(model.CardinalToAngle(mission.DropZone.Heading))
	camera.SetPitchAngle(0)
	camera.SetFovAngle(70, 1.0)

	// the map's static sprites, in the order tests/golden/make_fixtures.py freezes them
	sprites := make([]raycaster.Sprite, 0, 512)
	gold := make([]*goldSprite, 0, 512)
	for _, ms := range m.Sprites {
		img := tex.TextureImage(ms.Image)
		if img == nil {
			// not one of the map textures: load through the game's sprite loader
			img = resources.GetSpriteFromFile(ms.Image)
		}
		for _, p := range ms.Positions {
			g := &goldSprite{pos: geom.Vector2{X: p[0], Y: p[1]}, z: ms.ZPosition, scale: ms.Height / model.METERS_PER_UNIT,
				anchor: ms.Anchor.SpriteAnchor, img: img}
			gold = append(gold, g)
			sprites = append(sprites, g)
		}
	}

	camera.Update(sprites) // game/scene_game.go:161

	// TextureAt calls arrive from one goroutine per level: sort level-major, keeping the per-level order
	sort.SliceStable(rec.tex, func(i, j int) bool { return rec.tex[i].level < rec.tex[j].level })

	f, err := os.Create("golden_c1.bin")
	if err != nil {
		log.Fatal(err)
	}
	defer f.Close()
	le := binary.LittleEndian
	put := func(v any) {
		if err := binary.Write(f, le, v); err != nil {
			log.Fatal(err)
		}
	}
	f.Write([]byte("PM3DGOLD"))
	put([]int32{w, h, int32(m.NumLevels()), int32(len(rec.tex)), int32(len(rec.floor)), int32(len(gold))})
	for _, c := range rec.tex {
		put([]int32{c.x, c.y, c.level, c.side})
	}
	for _, c := range rec.floor {
		put([]int32{c.x, c.y})
	}
	for _, g := range gold {
		if g.rect == nil {
			put([]int32{0, 0, -1, -1})
		} else {
			put([]int32{int32(g.rect.Min.X), int32(g.rect.Min.Y), int32(g.rect.Max.X), int32(g.rect.Max.Y)})
		}
	}
	cp := camera.GetConvergencePoint()
	cx, cy := math.NaN(), math.NaN()
	if cp != nil {
		cx, cy = cp.X, cp.Y
	}
	put([]float64{camera.GetConvergenceDistance(), cx, cy})
	ci := int32(-1)
	if cs := camera.GetConvergenceSprite(); cs != nil {
		for i, g := range gold {
			if raycaster.Sprite(g) == cs {
				ci = int32(i)
			}
		}
	}
	put(ci)
	log.Printf("golden_c1.bin: %d TextureAt calls, %d FloorTextureAt calls, %d sprites", len(rec.tex), len(rec.floor), len(gold))
}
