// Package render — CUDA backend shim (drop this file into poly.red/render as render/cuda.go).
//
// cgo-free: libpolyred_cuda.so is bound through purego exactly like the reference's own GL backend
// (gpu/backend_gl_lib_linux.go:19-26, gpu/backend_gl.go:236-283): Dlopen + Dlsym + SyscallN, pointers
// passed as uintptr, floats only inside fixed-layout structs (include/polyred_cuda.h).
//
// NOT compiled in the build image of polyred-b200 (no Go toolchain there); the same logic is mirrored
// and tested in Python (polyred_b200/render.py), and tests/test_abi.py checks the struct layouts below
// against the C header with Go's alignment rules. See INTEGRATION.md for the reference-side patch (the hooks
// in options.go / raster.go and the one accessor in buffer/texture.go this file needs).
package render

import (
	"fmt"
	"image"
	"runtime"
	"unsafe"

	"github.com/ebitengine/purego"

	"poly.red/buffer"
	"poly.red/camera"
	"poly.red/color"
	"poly.red/geometry"
	"poly.red/geometry/primitive"
	"poly.red/internal/imageutil"
	"poly.red/light"
	"poly.red/material"
	"poly.red/math"
	"poly.red/scene"
	"poly.red/shader"
)

// CUDA selects the B200 backend for the whole render pass. It mirrors GPU(dev) (options.go:103-110); the
// renderer then never touches the CPU passes and never falls back: errors panic with the library message.
// Several devices = every frame is rendered by all of them (a device group inside the library, prc_group_*:
// raster passes partitioned by triangles, shading by screen strips, merged over NVLink) and is the one-device
// frame bit for bit; no devices = device 0.
func CUDA(devices ...int) Option {
	return func(o *option) {
		if len(devices) == 0 {
			devices = []int{0}
		}
		o.cudaDevices = append([]int(nil), devices...)
		o.useCUDA = true
		o.forceCPU = true // do not auto-open a Metal device (raster.go:110-122)
	}
}

const prcABIVersion = 1

const (
	prcMatFlat = 1 << iota
	prcMatAO
	prcMatRecvShadow
	prcMatNil
	prcMatNoMipmap
)

const (
	prcFramePerspect = 1 << iota
	prcFrameShadowMap
	prcFrameGamma
	prcFrameKeepGBuffer
	prcFrameNoReadback
	prcFrameUniformsResident
	prcFrameShadowReset
	prcFrameBGRA
	prcFrameAsync // with prcFrameNoReadback: enqueue only; prc_sync finishes (present path / view batches)
	prcFrameNoKernelTimers
	prcFrameImageAtSync // device groups set it themselves on asynchronous frames (include/polyred_cuda.h)
)

// ---- fixed-layout mirrors of include/polyred_cuda.h (little-endian, 8-byte aligned) ----

type prcMaterial struct {
	Diffuse, Specular uint32
	Shininess         float32
	Texture           int32
	Flags, _          uint32
}

type prcScene struct {
	ABIVersion, Flags       uint32
	NTris                   uint64
	Pos, Nor, UV            unsafe.Pointer // *float32
	Col                     unsafe.Pointer // *uint32
	Mat                     unsafe.Pointer // *int32
	NObjects, NMaterials    uint32
	ObjTriStart             unsafe.Pointer // *uint64
	Materials               unsafe.Pointer // *prcMaterial
	NTextures, NTexLevels   uint32
	TexFirstLevel           unsafe.Pointer // *uint32
	LevelW, LevelH          unsafe.Pointer // *uint32
	LevelOffset             unsafe.Pointer // *uint64
	TexData                 unsafe.Pointer // *uint8
	TexBytes                uint64
}

type prcObjectXf struct{ Trans, Normal [16]float32 }

type prcLight struct {
	Kind, CastShadow uint32
	Pos              [3]float32
	Intensity        float32
	Color, _         uint32
	View, Proj       [16]float32
	ShadowTrans      unsafe.Pointer // *[16]float32 per object
}

type prcFrame struct {
	ABIVersion, Flags, Width, Height      uint32
	NObjects, NLights, NAmbient, Backgrnd uint32
	Objects                               unsafe.Pointer // *prcObjectXf
	Lights                                unsafe.Pointer // *prcLight
	Ambient                               unsafe.Pointer // *float32
	Viewport, ViewportInv, ProjInv        [16]float32
	ViewInv, ViewportToWorld              [16]float32
	CamPos                                [3]float32
	MSAA                                  uint32 // render.MSAA(n); Width/Height above are n times render.Size (raster.go:149)
	GammaLUT                              [256]uint8
	Row0, Row1                            uint32
}

type cudaBackend struct {
	lib                                                                                               uintptr
	fnOpen, fnClose, fnLastError, fnSceneUpload, fnShadowReset, fnRender, fnHostImage, fnReadShadowmap uintptr
	fnGroupCtx                                                                                        uintptr
	ctx                                                                                               uintptr // prc_group*

	// what was flattened and uploaded: the scene and a signature of its membership (see sceneSignature)
	uploadedFor *scene.Scene
	uploadedSig uint64
	nObjects    int
}

func rgba(c color.RGBA) uint32 {
	return uint32(c.R) | uint32(c.G)<<8 | uint32(c.B)<<16 | uint32(c.A)<<24
}

func mat16(m math.Mat4[float32]) [16]float32 {
	return [16]float32{m.X00, m.X01, m.X02, m.X03, m.X10, m.X11, m.X12, m.X13, m.X20, m.X21, m.X22, m.X23, m.X30, m.X31, m.X32, m.X33}
}

func openCUDA(devices []int) *cudaBackend {
	lib, err := purego.Dlopen("libpolyred_cuda.so", purego.RTLD_NOW|purego.RTLD_GLOBAL)
	if err != nil {
		panic(fmt.Errorf("render: CUDA backend requested but libpolyred_cuda.so cannot be loaded: %w", err))
	}
	sym := func(name string) uintptr {
		p, err := purego.Dlsym(lib, name)
		if err != nil {
			panic(fmt.Errorf("render: libpolyred_cuda.so lacks %s: %w", name, err))
		}
		return p
	}
	// the group calls have the signatures of the single-context ones; a group of one device IS the single context
	b := &cudaBackend{lib: lib,
		fnOpen: sym("prc_group_open"), fnClose: sym("prc_group_close"), fnLastError: sym("prc_group_last_error"),
		fnSceneUpload: sym("prc_group_scene_upload"), fnShadowReset: sym("prc_group_shadow_reset"), fnRender: sym("prc_group_render"),
		fnHostImage: sym("prc_group_host_image"), fnGroupCtx: sym("prc_group_ctx"), fnReadShadowmap: sym("prc_read_shadowmap")}
	if v, _, _ := purego.SyscallN(sym("prc_abi_version")); uint32(v) != prcABIVersion {
		panic("render: libpolyred_cuda.so ABI version mismatch")
	}
	devs := make([]int32, len(devices))
	for i, d := range devices {
		devs[i] = int32(d)
	}
	rc, _, _ := purego.SyscallN(b.fnOpen, uintptr(unsafe.Pointer(unsafe.SliceData(devs))), uintptr(len(devs)), uintptr(unsafe.Pointer(&b.ctx)))
	runtime.KeepAlive(devs)
	if int32(rc) != 0 {
		panic(fmt.Errorf("render: prc_group_open(devices %v) failed with %d (no CPU fallback)", devices, int32(rc)))
	}
	runtime.SetFinalizer(b, func(b *cudaBackend) { purego.SyscallN(b.fnClose, b.ctx) })
	return b
}

// sceneSignature changes when the scene's membership does: which geometries it holds, in which order, with how many
// triangles and which materials (FNV-1a over pointers and counts). Model matrices are not part of it: the per-object
// uniforms are rebuilt from the scene graph every frame, so a moved TransformContext never needs a re-upload. Vertex data
// edited IN PLACE is not detected (as little as the reference detects it between NewRenderer and Render for its
// buffered meshes): call Options(Scene(s)) again to force the upload.
func sceneSignature(s *scene.Scene) uint64 {
	h := uint64(14695981039346656037)
	mix := func(v uint64) {
		for i := 0; i < 8; i++ {
			h ^= (v >> (8 * i)) & 0xff
			h *= 1099511628211
		}
	}
	scene.IterObjects(s, func(g *geometry.Geometry, _ math.Mat4[float32]) bool {
		mix(uint64(uintptr(unsafe.Pointer(g))))
		ts := g.Triangles()
		mix(uint64(len(ts)))
		if len(ts) > 0 {
			mix(uint64(uintptr(unsafe.Pointer(ts[0]))))
		}
		for _, m := range g.Materials() {
			if bp, _ := m.(*material.BlinnPhong); bp != nil {
				mix(uint64(uintptr(unsafe.Pointer(bp))))
				mix(uint64(uintptr(unsafe.Pointer(bp.Texture))))
			} else {
				mix(0)
			}
		}
		return true
	})
	return h
}

func (b *cudaBackend) check(rc uintptr, what string) {
	if int32(rc) == 0 {
		return
	}
	p, _, _ := purego.SyscallN(b.fnLastError, b.ctx)
	msg := ""
	for q := (*byte)(unsafe.Pointer(p)); q != nil && *q != 0; q = (*byte)(unsafe.Add(unsafe.Pointer(q), 1)) {
		msg += string(rune(*q))
	}
	panic(fmt.Errorf("render: %s failed (%d): %s", what, int32(rc), msg))
}

// uploadScene flattens the scene graph in draw order, exactly the walk cpuForwardPass does every frame
// (raster.go:241-270): flat material id = base + local, negative ids stay negative.
func (b *cudaBackend) uploadScene(s *scene.Scene) {
	var (
		pos, nor, uv []float32
		col          []uint32
		mat          []int32
		objStart     = []uint64{0}
		mats         []prcMaterial
		texIndex     = map[*material.BlinnPhong]int32{}
		texFirst     = []uint32{0}
		levelW       []uint32
		levelH       []uint32
		levelOff     []uint64
		texData      []uint8
	)
	scene.IterObjects(s, func(g *geometry.Geometry, _ math.Mat4[float32]) bool {
		base := int32(len(mats))
		for _, m := range g.Materials() {
			bp, _ := m.(*material.BlinnPhong)
			pm := prcMaterial{Texture: -1, Flags: prcMatNil}
			if bp != nil && bp.Texture == nil {
				// FragmentShader dereferences m.Texture (shader/blinn_cpu.go:30-36): the reference panics there. No defined
				// result => an error, never a silent change of appearance.
				panic("render: BlinnPhong material without a texture (the CPU path nil-dereferences it, shader/blinn_cpu.go:30)")
			}
			if bp != nil {
				pm = prcMaterial{Diffuse: rgba(bp.Diffuse), Specular: rgba(bp.Specular), Shininess: bp.Shininess}
				if bp.FlatShading {
					pm.Flags |= prcMatFlat
				}
				if bp.AmbientOcclusion {
					pm.Flags |= prcMatAO
				}
				if bp.ReceiveShadow {
					pm.Flags |= prcMatRecvShadow
				}
				if !bp.Texture.UseMipmap() {
					pm.Flags |= prcMatNoMipmap
				}
				idx, ok := texIndex[bp]
				if !ok {
					idx = int32(len(texFirst) - 1)
					texIndex[bp] = idx
					// Texture.Mipmaps() is the one accessor this shim needs added to buffer.Texture
					// (it returns t.mipmap, buffer/texture.go:31).
					for _, lv := range bp.Texture.Mipmaps() {
						levelW = append(levelW, uint32(lv.Bounds().Dx()))
						levelH = append(levelH, uint32(lv.Bounds().Dy()))
						levelOff = append(levelOff, uint64(len(texData)))
						for y := 0; y < lv.Bounds().Dy(); y++ {
							texData = append(texData, lv.Pix[y*lv.Stride:y*lv.Stride+4*lv.Bounds().Dx()]...)
						}
					}
					texFirst = append(texFirst, uint32(len(levelW)))
				}
				pm.Texture = idx
			}
			mats = append(mats, pm)
		}
		for _, t := range g.Triangles() {
			for _, v := range [3]*primitive.Vertex{t.V1, t.V2, t.V3} {
				if v.Pos.W != 1 {
					panic("render: CUDA backend requires Pos.W == 1 (what every loader produces, model/load.go:162-164)")
				}
				pos = append(pos, v.Pos.X, v.Pos.Y, v.Pos.Z)
				nor = append(nor, v.Nor.X, v.Nor.Y, v.Nor.Z)
				uv = append(uv, v.UV.X, v.UV.Y)
				col = append(col, rgba(v.Col))
			}
			id := int32(t.MaterialID)
			if id >= 0 {
				id += base
			}
			mat = append(mat, id)
		}
		objStart = append(objStart, uint64(len(mat)))
		return true
	})
	sc := prcScene{ABIVersion: prcABIVersion, NTris: uint64(len(mat)),
		Pos: unsafe.Pointer(unsafe.SliceData(pos)), Nor: unsafe.Pointer(unsafe.SliceData(nor)), UV: unsafe.Pointer(unsafe.SliceData(uv)),
		Col: unsafe.Pointer(unsafe.SliceData(col)), Mat: unsafe.Pointer(unsafe.SliceData(mat)),
		NObjects: uint32(len(objStart) - 1), NMaterials: uint32(len(mats)),
		ObjTriStart: unsafe.Pointer(unsafe.SliceData(objStart)), Materials: unsafe.Pointer(unsafe.SliceData(mats)),
		NTextures: uint32(len(texFirst) - 1), NTexLevels: uint32(len(levelW)),
		TexFirstLevel: unsafe.Pointer(unsafe.SliceData(texFirst)), LevelW: unsafe.Pointer(unsafe.SliceData(levelW)),
		LevelH: unsafe.Pointer(unsafe.SliceData(levelH)), LevelOffset: unsafe.Pointer(unsafe.SliceData(levelOff)),
		TexData: unsafe.Pointer(unsafe.SliceData(texData)), TexBytes: uint64(len(texData))}
	rc, _, _ := purego.SyscallN(b.fnSceneUpload, b.ctx, uintptr(unsafe.Pointer(&sc)))
	runtime.KeepAlive(pos); runtime.KeepAlive(nor); runtime.KeepAlive(uv); runtime.KeepAlive(col); runtime.KeepAlive(mat)
	runtime.KeepAlive(objStart); runtime.KeepAlive(mats); runtime.KeepAlive(texFirst); runtime.KeepAlive(levelW)
	runtime.KeepAlive(levelH); runtime.KeepAlive(levelOff); runtime.KeepAlive(texData)
	b.check(rc, "prc_group_scene_upload")
	b.uploadedFor, b.uploadedSig, b.nObjects = s, sceneSignature(s), len(objStart)-1
}

// renderCUDA is (*Renderer).Render() for the CUDA backend (raster.go:155-199): it computes the same
// uniforms as cpuForwardPass (raster.go:232-246), passDeferred (raster.go:281-295) and passShadows
// (shadow.go:121-135) with the reference's own math package, then makes ONE library call.
func (r *Renderer) renderCUDA() *image.RGBA {
	b := r.cuda
	if b == nil {
		b = openCUDA(r.cfg.cudaDevices)
		r.cuda = b
	}
	if r.cfg.BlendFunc != nil {
		panic("render: CUDA backend does not support Blending")
	}
	if b.uploadedFor != r.cfg.Scene || b.uploadedSig != sceneSignature(r.cfg.Scene) {
		b.uploadScene(r.cfg.Scene) // a new scene, or objects were added / removed / replaced (SURVEY 8f-2)
	}
	if r.cudaShadowDirty { // set by initShadowMaps (NewRenderer / Options): fresh zero maps (shadow.go:87)
		rc, _, _ := purego.SyscallN(b.fnShadowReset, b.ctx)
		b.check(rc, "prc_group_shadow_reset")
		r.cudaShadowDirty = false
	}
	// the frame buffer is MSAA times render.Size (resetBufs, raster.go:149); the library resizes back (raster.go:377)
	w, h := r.cfg.Width*r.cfg.MSAA, r.cfg.Height*r.cfg.MSAA
	view, proj := r.cfg.Camera.ViewMatrix(), r.cfg.Camera.ProjMatrix()
	vp := math.ViewportMatrix(float32(w), float32(h))
	viewInv, projInv, vpInv := view.Inv(), proj.Inv(), vp.Inv()

	var models []math.Mat4[float32]
	var objs []prcObjectXf
	scene.IterObjects(r.cfg.Scene, func(g *geometry.Geometry, modelMatrix math.Mat4[float32]) bool {
		mvp := shader.MVP{Model: modelMatrix.MulM(g.ModelMatrix())}
		mvp.Normal = mvp.Model.Inv().T()
		models = append(models, mvp.Model)
		objs = append(objs, prcObjectXf{Trans: mat16(proj.MulM(view).MulM(mvp.Model)), Normal: mat16(mvp.Normal)})
		return true
	})
	ls, es := r.cfg.Scene.Lights()
	lights := make([]prcLight, len(ls))
	shadowTrans := make([][][16]float32, len(ls))
	for i, l := range ls {
		pl := &lights[i]
		switch ll := l.(type) {
		case *light.Point:
			p := ll.Position()
			pl.Kind, pl.Pos = 0, [3]float32{p.X, p.Y, p.Z}
		case *light.Directional:
			d := ll.Dir()
			pl.Kind, pl.Pos = 1, [3]float32{d.X, d.Y, d.Z}
		default:
			panic("render: CUDA backend supports Point and Directional sources (shader/blinn_cpu.go:72-80)")
		}
		pl.Intensity, pl.Color = l.Intensity(), rgba(l.Color())
		if r.cfg.ShadowMap && l.CastShadow() {
			var cam camera.Interface = r.shadowBufs[i].camera
			if cam == nil {
				panic("render: shadow-casting light without a light camera (the CPU path panics too, shadow.go:123)")
			}
			lv, lp := cam.ViewMatrix(), cam.ProjMatrix()
			pl.CastShadow, pl.View, pl.Proj = 1, mat16(lv), mat16(lp)
			st := make([][16]float32, len(models))
			for o := range models {
				st[o] = mat16(lp.MulM(lv).MulM(models[o])) // shadow.go:155
			}
			shadowTrans[i] = st
			pl.ShadowTrans = unsafe.Pointer(unsafe.SliceData(st))
		}
	}
	amb := make([]float32, len(es))
	for i, e := range es {
		amb[i] = e.Intensity()
	}
	f := prcFrame{ABIVersion: prcABIVersion, Width: uint32(w), Height: uint32(h),
		NObjects: uint32(len(objs)), NLights: uint32(len(ls)), NAmbient: uint32(len(es)), Backgrnd: rgba(r.cfg.Background),
		Objects: unsafe.Pointer(unsafe.SliceData(objs)), Lights: unsafe.Pointer(unsafe.SliceData(lights)), Ambient: unsafe.Pointer(unsafe.SliceData(amb)),
		Viewport: mat16(vp), ViewportInv: mat16(vpInv), ProjInv: mat16(projInv), ViewInv: mat16(viewInv),
		ViewportToWorld: mat16(viewInv.MulM(projInv).MulM(vpInv)), Row0: 0, Row1: uint32(h), MSAA: uint32(r.cfg.MSAA)}
	c := r.cfg.Camera.Position()
	f.CamPos = [3]float32{c.X, c.Y, c.Z}
	if r.cfg.Perspect {
		f.Flags |= prcFramePerspect
	}
	if r.cfg.ShadowMap {
		f.Flags |= prcFrameShadowMap
	}
	if r.cfg.GammaCorrect {
		f.Flags |= prcFrameGamma
	}
	if r.cfg.Format == buffer.PixelFormatBGRA {
		f.Flags |= prcFrameBGRA
	}
	for i := 0; i < 256; i++ { // shader.GammaCorrection (shader/gamma.go:13-18) as a table
		f.GammaLUT[i] = uint8(color.FromLinear2sRGB(float32(i)/0xff)*0xff + 0.5)
	}
	// rgba_out == NULL: the frame lands in the library's page-locked double buffer and is wrapped in place
	// (zero copy). Like the reference's own double buffer (raster.go:86,201-206) it is valid until two frames later.
	rc, _, _ := purego.SyscallN(b.fnRender, b.ctx, uintptr(unsafe.Pointer(&f)), 0)
	runtime.KeepAlive(objs); runtime.KeepAlive(lights); runtime.KeepAlive(amb); runtime.KeepAlive(shadowTrans)
	b.check(rc, "prc_group_render")
	if r.cfg.Debug && r.cfg.ShadowMap {
		// render.Debug(true): passShadows saves shadow-<i>.png from shadowInfo.depths (shadow.go:98-118). The maps live in HBM;
		// they are read back into the renderer's own depth slices (same size and index order, shadow.go:26-31,221-228) and
		// dumped with the reference's loop.
		for i, l := range ls {
			if !l.CastShadow() {
				continue
			}
			var ctx0 uintptr // every rank of a group holds the merged maps: read rank 0's
			rc, _, _ := purego.SyscallN(b.fnGroupCtx, b.ctx, 0, uintptr(unsafe.Pointer(&ctx0)))
			b.check(rc, "prc_group_ctx")
			rc, _, _ = purego.SyscallN(b.fnReadShadowmap, ctx0, uintptr(i), uintptr(unsafe.Pointer(unsafe.SliceData(r.shadowBufs[i].depths))))
			b.check(rc, "prc_read_shadowmap")
			img := image.NewRGBA(image.Rect(0, 0, w, h))
			for x := 0; x < w; x++ {
				for y := 0; y < h; y++ {
					z := uint8(r.shadowBufs[i].depths[x+(h-y-1)*w] * 255)
					img.SetRGBA(x, y, color.RGBA{z, z, z, 255})
				}
			}
			file := fmt.Sprintf("shadow-%d.png", i)
			fmt.Printf("saving (shadow map)... %s\n", file)
			imageutil.Save(img, file)
		}
	}
	var ptr, n uint64
	rc, _, _ = purego.SyscallN(b.fnHostImage, b.ctx, uintptr(unsafe.Pointer(&ptr)), uintptr(unsafe.Pointer(&n)))
	b.check(rc, "prc_group_host_image")
	ow, oh := r.cfg.Width, r.cfg.Height
	r.outBuf = &image.RGBA{Pix: unsafe.Slice((*uint8)(unsafe.Pointer(uintptr(ptr))), int(n)), Stride: 4 * ow, Rect: image.Rect(0, 0, ow, oh)}
	r.passGPU["forward"], r.passGPU["deferred"], r.passGPU["gamma"] = true, true, true
	return r.outBuf
}
