"""The pybind11 host layer (`import psdr_cuda`): CPU tests of its C++ scene ingest against the oracle's loader and of the
reference-facing class surface; GPU tests of renderC / renderD + torch.autograd against the ctypes path."""
import numpy as np
import pytest

from conftest import scene_path


@pytest.fixture(scope="module")
def psdr_cuda(native_lib):
    import psdr_cuda_b200.compat  # noqa: F401
    import psdr_cuda as m
    return m


def test_surface_matches_reference_names(psdr_cuda):
    # src/psdr.cpp:48-294
    for name in ("Scene", "RenderOption", "Mesh", "DiffuseBSDF", "RoughConductorBSDF", "PerspectiveCamera", "AreaLight", "EnvironmentMap",
                 "Bitmap1fD", "Bitmap3fD", "Integrator", "DirectIntegrator", "FieldExtractionIntegrator", "PathIntegrator"):
        assert hasattr(psdr_cuda, name), name
    o = psdr_cuda.RenderOption(64, 48, 4)
    assert (o.width, o.height, o.spp, o.sppe, o.sppse) == (64, 48, 4, 4, 4)
    o = psdr_cuda.RenderOption(64, 48, 4, 2, 1)
    assert (o.sppe, o.sppse) == (2, 1)
    o.spp = 9
    assert o.spp == 9
    d = psdr_cuda.DirectIntegrator(2, 3)
    d.hide_emitters = True
    assert d.hide_emitters
    with pytest.raises(RuntimeError):
        psdr_cuda.DirectIntegrator(0, 0)
    with pytest.raises(RuntimeError):
        psdr_cuda.FieldExtractionIntegrator("nonsense")


@pytest.mark.parametrize("name", ["cbox_bunny", "cbox_bunny_mutiemitter", "bunny", "tree", "bunny_env", "bunny_env_2"])
def test_cpp_loader_matches_oracle_loader(psdr_cuda, name):
    from oracle import orc
    ref = orc.load_scene_description(scene_path(name))
    sc = psdr_cuda.Scene(-1)                    # description only: no CUDA context needed for ingest
    sc.load_file(scene_path(name), False)
    assert sc.num_meshes == len(ref["meshes"]) and sc.num_sensors == len(ref["sensors"])
    assert (sc.opts.width, sc.opts.height, sc.opts.spp, sc.opts.sppe, sc.opts.sppse) == tuple(ref["opts"][k] for k in ("width", "height", "spp", "sppe", "sppse"))
    pm = sc.param_map
    for i, m in enumerate(ref["meshes"]):
        o = pm["Mesh[%d]" % i]
        assert o.type_name() == "Mesh" and o.num_vertices == len(m["verts"]) and o.num_faces == len(m["faces"])
        assert np.array_equal(o.vertex_positions, m["verts"]) and np.array_equal(o.face_indices, m["faces"])
        assert np.array_equal(o.to_world_raw, m["to_world"]) and o.bsdf_index == m["bsdf"] and o.use_face_normals == m["face_normals"]
    for i, b in enumerate(ref["bsdfs"]):
        o = pm["BSDF[%d]" % i]
        assert pm["BSDF[id=%s]" % b["id"]].id == b["id"]
        if b["type"] == 0:
            assert o.type_name() == "Diffuse" and np.array_equal(o.reflectance.data.reshape(-1), b["reflectance"].reshape(-1))
        else:
            assert o.type_name() == "RoughConductor" and np.array_equal(o.alpha_u.data.reshape(-1), b["alpha_u"].reshape(-1))
            assert np.array_equal(o.eta.data.reshape(-1), b["eta"].reshape(-1)) and np.array_equal(o.k.data.reshape(-1), b["k"].reshape(-1))
    for i, s in enumerate(ref["sensors"]):
        o = pm["Sensor[%d]" % i]
        assert np.array_equal(o.to_world, s["to_world"]) and o.fov_x == np.float32(s["fov"])
    if ref["envmap"] is not None:
        e = pm["Emitter[0]"]
        assert e.type_name() == "AreaLight"      # sic, include/psdr/emitter/envmap.h:58
        assert np.array_equal(e.radiance.data.reshape(ref["envmap"]["radiance"].shape), ref["envmap"]["radiance"])
        assert e.scale == np.float32(ref["envmap"]["scale"])


def test_loader_error_messages(psdr_cuda):
    sc = psdr_cuda.Scene(-1)
    with pytest.raises(RuntimeError, match="XML parsing failed"):
        sc.load_string("<scene", False)
    with pytest.raises(RuntimeError, match="Unsupported BSDF"):
        psdr_cuda.Scene(-1).load_string("<scene><bsdf type='plastic' id='x'/></scene>", False)
    with pytest.raises(RuntimeError, match="BSDF must have an id"):
        psdr_cuda.Scene(-1).load_string("<scene><bsdf type='diffuse'><rgb name='reflectance' value='1'/></bsdf></scene>", False)
    sc = psdr_cuda.Scene(-1)
    sc.load_file(scene_path("bunny"), False)
    with pytest.raises(RuntimeError, match="already loaded"):
        sc.load_file(scene_path("bunny"), False)
    with pytest.raises(RuntimeError, match="CUDA device"):
        sc.configure()                           # no context, no CPU fallback
    with pytest.raises(RuntimeError, match="must be configured"):
        psdr_cuda.DirectIntegrator(1, 1).renderC_numpy(sc, 0)


def test_mesh_transform_semantics(psdr_cuda):
    # mesh.h:19-35: set_transform overwrites left (default), append_transform pre-multiplies left / post-multiplies right
    sc = psdr_cuda.Scene(-1)
    sc.load_file(scene_path("bunny"), False)
    m = sc.param_map["Mesh[0]"]
    raw = m.to_world_raw
    T = np.eye(4, dtype=np.float32); T[0, 3] = 2
    S = np.diag(np.array([2, 2, 2, 1], np.float32))
    m.set_transform(T)
    assert np.allclose(m.to_world, T @ raw)
    m.append_transform(S)
    assert np.allclose(m.to_world, S @ T @ raw)
    m.append_transform(S, False)
    assert np.allclose(m.to_world, S @ T @ raw @ S)


@pytest.mark.gpu
def test_psdr_cuda_module_renders_and_differentiates(psdr_cuda):
    torch = pytest.importorskip("torch")
    from psdr_cuda_b200 import capi, scene_io
    sc = psdr_cuda.Scene()
    sc.load_file(scene_path("cbox_bunny"), False)
    sc.opts.width, sc.opts.height, sc.opts.spp, sc.opts.sppe, sc.opts.sppse = 48, 48, 4, 4, 4
    albedo = sc.parameter("BSDF[id=white]", "reflectance")
    verts = sc.parameter("Mesh[1]", "vertex_positions")
    sc.configure()
    integ = psdr_cuda.PathIntegrator(2)
    img_c = integ.renderC(sc, 0)
    img = integ.renderD(sc, 0)
    assert img.shape == (48 * 48, 3) and img.requires_grad
    w = torch.linspace(0.5, 1.5, img.numel(), device=img.device).view_as(img)
    (img * w).sum().backward()
    assert albedo.grad is not None and verts.grad is not None and verts.grad.shape == (34817, 3)
    # same thing through the ctypes binding
    ctx = capi.Context(0)
    ctx.load_description(scene_io.load_scene_description(scene_path("cbox_bunny")), dict(width=48, height=48, spp=4, sppe=4, sppse=4))
    ctx.grad_require(capi.PARAM_BSDF_TEXTURE, 0, "reflectance")
    ctx.grad_require(capi.PARAM_MESH_VERTICES, 1)
    ctx.configure()
    ci = capi.make_integrator("path", max_depth=2)
    ref_c = ctx.render_c(ci)
    ref_d = ctx.render_d(ci)
    g = ctx.render_d_vjp(ci, w.contiguous())
    assert torch.equal(img_c, ref_c) and torch.equal(img.detach(), ref_d)
    assert torch.allclose(albedo.grad.reshape(-1), g[:3], rtol=1e-4, atol=1e-5)
    assert (verts.grad.reshape(-1) - g[3:]).norm() <= 1e-4 * g[3:].norm()
    # forward mode through the module: directional derivative of the image along d(albedo) = (1, 1, 1)
    img_f, dimg = integ.forward(sc, {albedo: torch.ones_like(albedo)})
    ctx.render_d(ci)
    flat = torch.zeros(ctx.grad_size(), device="cuda"); flat[:3] = 1
    assert torch.allclose(dimg, ctx.render_d_jvp(ci, flat), rtol=1e-5, atol=1e-6)
    # an optimisation step moves the parameter and the next configure() picks it up (examples/utils/adam.py flow)
    with torch.no_grad():
        albedo -= 0.1 * albedo.grad / albedo.grad.abs().max()
    sc.configure()
    assert not torch.equal(integ.renderC(sc, 0), img_c)


@pytest.mark.gpu
def test_mesh_transform_leaf_gradient_matches_oracle(psdr_cuda):
    """Mesh.set_transform as a differentiable leaf (src/psdr.cpp:246-247, examples/utils/differential.py:7-11): reverse mode
    through the module against the oracle's forward mode for a translation and a rotation-like tangent of to_world_left."""
    torch = pytest.importorskip("torch")
    from oracle import orc
    sc = psdr_cuda.Scene()
    sc.load_file(scene_path("cbox_bunny"), False)
    sc.opts.width, sc.opts.height, sc.opts.spp, sc.opts.sppe, sc.opts.sppse = 40, 40, 4, 4, 4
    T = sc.parameter("Mesh[1]", "to_world_left")
    sc.configure()
    integ = psdr_cuda.DirectIntegrator(1, 1)
    img = integ.renderD(sc, 0)
    rng = np.random.default_rng(9)
    w = torch.from_numpy(rng.uniform(-1, 1, size=(40 * 40, 3)).astype(np.float32)).to(img.device)
    (img * w).sum().backward()
    assert T.grad is not None and T.grad.shape == (4, 4)
    g = T.grad.double().cpu().numpy()
    odesc = orc.load_scene_description(scene_path("cbox_bunny"))
    opts = dict(width=40, height=40, spp=4, sppe=4, sppse=4)
    tangents = [np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32)]
    tangents[0][0, 3] = 1.0                                            # d/dP translate(P, 0, 0)
    tangents[1][0, 1], tangents[1][1, 0], tangents[1][2, 3] = -1.0, 1.0, 0.5   # d/dP rotate(z, P) at P = 0, plus a z shift
    for dM in tangents:
        osc = orc.Scene(odesc, opts)
        osc.set_mesh_transform_tangent(1, dM, True)
        osc.configure()
        _, dimg = orc.DirectIntegrator(1, 1).renderD(osc)
        want = float((w.double().cpu().numpy() * dimg).sum())
        got = float((g * dM).sum())
        assert abs(got - want) <= 3e-3 * abs(want), (got, want)


@pytest.mark.gpu
def test_roughconductor_and_envmap_leaves_through_the_module(psdr_cuda):
    """RoughConductorBSDF.{alpha_u,eta}.data, EnvironmentMap.{radiance.data,scale} and Mesh.vertex_positions as torch leaves
    (src/psdr.cpp:211-215,236-238): one backward through the module equals the VJP of the ctypes binding."""
    torch = pytest.importorskip("torch")
    from psdr_cuda_b200 import capi, scene_io
    sc = psdr_cuda.Scene()
    sc.load_file(scene_path("bunny_env"), False)
    sc.opts.width, sc.opts.height, sc.opts.spp, sc.opts.sppe, sc.opts.sppse = 32, 32, 4, 0, 0
    alpha = sc.parameter("BSDF[0]", "alpha_u")
    eta = sc.parameter("BSDF[id=mat1]", "eta")
    rad = sc.parameter("Emitter[0]", "radiance")
    scale = sc.parameter("Emitter[0]", "scale")
    verts = sc.parameter("Mesh[0]", "vertex_positions")
    sc.configure()
    integ = psdr_cuda.DirectIntegrator(1, 1)
    img = integ.renderD(sc, 0)
    w = torch.linspace(-1.0, 1.0, img.numel(), device=img.device).view_as(img)
    (img * w).sum().backward()
    for t in (alpha, eta, rad, scale, verts):
        assert t.grad is not None and torch.isfinite(t.grad).all()
    ctx = capi.Context(0)
    ctx.load_description(scene_io.load_scene_description(scene_path("bunny_env")), dict(width=32, height=32, spp=4, sppe=0, sppse=0))
    ctx.grad_require(capi.PARAM_BSDF_TEXTURE, 0, "alpha_u")
    ctx.grad_require(capi.PARAM_BSDF_TEXTURE, 0, "eta")
    ctx.grad_require(capi.PARAM_MESH_VERTICES, 0)
    ctx.grad_require(capi.PARAM_ENVMAP_RADIANCE, 0)
    ctx.grad_require(capi.PARAM_ENVMAP_SCALE, 0)
    ctx.configure()
    ci = capi.make_integrator("direct", bsdf_samples=1, light_samples=1)
    ref = ctx.render_d(ci)
    assert torch.equal(img.detach(), ref)
    g = ctx.render_d_vjp(ci, w.contiguous())
    seg = {(s["kind"], s["slot"]): s for s in ctx.grad_layout()}
    def part(kind, slot=0):
        s = seg[(kind, slot)]
        return g[s["offset"]:s["offset"] + s["count"]]
    def close(a, b):
        return (a.reshape(-1) - b).norm() <= 1e-4 * b.norm() + 1e-7
    assert close(alpha.grad, part(capi.PARAM_BSDF_TEXTURE, capi.TEX["alpha_u"]))
    assert close(eta.grad, part(capi.PARAM_BSDF_TEXTURE, capi.TEX["eta"]))
    assert close(verts.grad, part(capi.PARAM_MESH_VERTICES))
    assert close(rad.grad, part(capi.PARAM_ENVMAP_RADIANCE))
    assert close(scale.grad, part(capi.PARAM_ENVMAP_SCALE))
    # editing the leaves reaches the renderer at the next configure()
    with torch.no_grad():
        scale *= 0.5
    sc.configure()
    img2 = integ.renderC(sc, 0)
    assert abs(float(img2.mean()) / float(integ.renderC(sc, 0).mean()) - 1.0) < 0.2 and float(img2.mean()) < 0.75 * float(img.detach().mean())


@pytest.mark.gpu
def test_cuda_leaves_update_the_scene_without_leaving_the_device(psdr_cuda):
    """after the first configure, registered torch leaves reach the scene device-to-device (pb_scene_set_*_device): the optimisation loop of
    docs/inverse_diff_render.rst keeps its parameters on the GPU like the reference's Enoki arrays; host mirrors refresh when read"""
    torch = pytest.importorskip("torch")
    sc = psdr_cuda.Scene()
    sc.load_file(scene_path("cbox_bunny"), False)
    sc.opts.width, sc.opts.height, sc.opts.spp, sc.opts.sppe, sc.opts.sppse, sc.opts.log_level = 32, 32, 4, 0, 0, 0
    verts = sc.parameter("Mesh[1]", "vertex_positions")
    albedo = sc.parameter("BSDF[id=red]", "reflectance")
    sc.configure()
    integ = psdr_cuda.PathIntegrator(2)
    before = integ.renderC(sc, 0)
    with torch.no_grad():
        verts += torch.tensor([3.0, 0.0, 0.0], device=verts.device)
        albedo.copy_(torch.tensor([[0.1, 0.8, 0.1]], device=albedo.device))
    sc.configure()                                         # device-to-device
    after = integ.renderC(sc, 0)
    assert not torch.equal(before, after)
    host = np.asarray(sc.param_map["Mesh[1]"].vertex_positions)
    assert np.allclose(host, verts.detach().cpu().numpy())
    # the same edit through the host path gives the same image
    sc2 = psdr_cuda.Scene()
    sc2.load_file(scene_path("cbox_bunny"), False)
    sc2.opts.width, sc2.opts.height, sc2.opts.spp, sc2.opts.sppe, sc2.opts.sppse, sc2.opts.log_level = 32, 32, 4, 0, 0, 0
    sc2.param_map["Mesh[1]"].vertex_positions = verts.detach().cpu().numpy()
    sc2.param_map["BSDF[id=red]"].reflectance.data = np.array([[0.1, 0.8, 0.1]], np.float32)
    sc2.configure()
    integ.renderC(sc2, 0)                                  # same position of the sampler streams as `after`
    assert torch.allclose(integ.renderC(sc2, 0), after, atol=1e-6)


@pytest.mark.gpu
def test_two_renders_before_one_backward_use_their_own_samples(psdr_cuda):
    """A multi-view loss renders several images before one backward (the reference's tape differentiates each with the samples
    that produced it). Every renderD node remembers its sampler positions (pb_render_d_get_state) and restores them for its VJP:
    the gradient of view A must not change because view B was rendered in between."""
    torch = pytest.importorskip("torch")
    sc = psdr_cuda.Scene()
    sc.load_file(scene_path("cbox_bunny"), False)
    sc.opts.width, sc.opts.height, sc.opts.spp, sc.opts.sppe, sc.opts.sppse, sc.opts.log_level = 48, 48, 4, 2, 2, 0
    albedo = sc.parameter("BSDF[id=white]", "reflectance")
    verts = sc.parameter("Mesh[1]", "vertex_positions")
    sc.configure()
    integ = psdr_cuda.PathIntegrator(2)
    w = torch.from_numpy(np.random.default_rng(5).uniform(-1, 1, size=(48 * 48, 3)).astype(np.float32)).cuda()
    # A alone
    img_a = integ.renderD(sc, 0)
    (img_a * w).sum().backward()
    ga, gv = albedo.grad.clone(), verts.grad.clone()
    albedo.grad = None; verts.grad = None
    # the same A, then B, then one backward through A only / through both
    sc2 = psdr_cuda.Scene()                             # a fresh scene starts the sampler streams again
    sc2.load_file(scene_path("cbox_bunny"), False)
    sc2.opts.width, sc2.opts.height, sc2.opts.spp, sc2.opts.sppe, sc2.opts.sppse, sc2.opts.log_level = 48, 48, 4, 2, 2, 0
    albedo2 = sc2.parameter("BSDF[id=white]", "reflectance")
    verts2 = sc2.parameter("Mesh[1]", "vertex_positions")
    sc2.configure()
    img_a2 = integ.renderD(sc2, 0)
    img_b2 = integ.renderD(sc2, 0)                      # a second view (same sensor, next samples) before backward
    assert torch.equal(img_a2.detach(), img_a.detach()) and not torch.equal(img_b2.detach(), img_a2.detach())
    (img_a2 * w).sum().backward(retain_graph=True)
    assert torch.allclose(albedo2.grad, ga, rtol=1e-4, atol=1e-6)
    assert (verts2.grad - gv).norm() <= 1e-3 * gv.norm()
    albedo2.grad = None; verts2.grad = None
    ((img_a2 + img_b2) * w).sum().backward()            # both nodes' VJPs, each with its own samples
    sc3 = psdr_cuda.Scene()
    sc3.load_file(scene_path("cbox_bunny"), False)
    sc3.opts.width, sc3.opts.height, sc3.opts.spp, sc3.opts.sppe, sc3.opts.sppse, sc3.opts.log_level = 48, 48, 4, 2, 2, 0
    albedo3 = sc3.parameter("BSDF[id=white]", "reflectance")
    sc3.configure()
    integ.renderD(sc3, 0)
    img_b3 = integ.renderD(sc3, 0)
    (img_b3 * w).sum().backward()                       # B alone, rendered last: the plain path
    assert torch.allclose(albedo2.grad, ga + albedo3.grad, rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
def test_inverse_rendering_loop_recovers_albedo_and_translation(psdr_cuda):
    """The use the module exists for (docs/inverse_diff_render.rst:48-79, examples/utils/adam.py): gradient descent through
    renderD on a material parameter and on a mesh transform (boundary terms on, BVH refit between iterations)."""
    torch = pytest.importorskip("torch")

    def scene(spp, sppe, sppse):
        sc = psdr_cuda.Scene()
        sc.load_file(scene_path("cbox_bunny"), False)
        sc.opts.width, sc.opts.height, sc.opts.spp, sc.opts.sppe, sc.opts.sppse = 64, 64, spp, sppe, sppse
        return sc

    integ = psdr_cuda.DirectIntegrator(1, 1)
    # target: the fixture as it is
    ref = scene(32, 0, 0); ref.configure()
    target = integ.renderC(ref, 0).clone()

    # (1) red wall albedo from a wrong start
    sc = scene(16, 0, 0)
    albedo = sc.parameter("BSDF[id=red]", "reflectance")
    truth = albedo.detach().clone()
    with torch.no_grad():
        albedo.copy_(torch.tensor([0.3, 0.6, 0.6], device=albedo.device).view_as(albedo))
    opt = torch.optim.Adam([albedo], lr=0.05)
    err0 = float((albedo.detach() - truth).abs().max())
    for it in range(40):
        sc.configure()
        opt.zero_grad()
        loss = (integ.renderD(sc, 0) - target).square().mean()
        loss.backward()
        opt.step()
        with torch.no_grad():
            albedo.clamp_(0.0, 1.0)
    assert float((albedo.detach() - truth).abs().max()) < 0.25 * err0

    # (2) bunny translated along x: recover the offset through the primary / secondary boundary terms
    sc = scene(8, 8, 8)
    T = sc.parameter("Mesh[1]", "to_world_left")
    with torch.no_grad():
        T[0, 3] = 12.0
    opt = torch.optim.Adam([T], lr=1.0)
    losses = []
    for it in range(30):
        sc.configure()
        opt.zero_grad()
        loss = (integ.renderD(sc, 0) - target).square().mean()
        loss.backward()
        with torch.no_grad():   # a translation along x is the only degree of freedom of this test
            g = T.grad[0, 3].clone(); T.grad.zero_(); T.grad[0, 3] = g
        opt.step()
        losses.append(float(loss.detach()))
    assert abs(float(T.detach()[0, 3])) < 4.0, (float(T.detach()[0, 3]), losses[0], losses[-1])
    assert losses[-1] < 0.85 * losses[0]   # the rest is Monte-Carlo noise of 8 spp against the 32 spp target


@pytest.mark.gpu
def test_sensor_pose_leaf_through_the_module(psdr_cuda):
    """Sensor.to_world as a torch leaf (src/psdr.cpp:220-224): backward through the module equals the ctypes VJP; a pose
    optimisation step moves the image towards the target."""
    torch = pytest.importorskip("torch")
    from psdr_cuda_b200 import capi, scene_io
    sc = psdr_cuda.Scene()
    sc.load_file(scene_path("cbox_bunny"), False)
    sc.opts.width, sc.opts.height, sc.opts.spp, sc.opts.sppe, sc.opts.sppse = 40, 40, 4, 4, 4
    pose = sc.parameter("Sensor[0]", "to_world")
    sc.configure()
    integ = psdr_cuda.DirectIntegrator(1, 1)
    img = integ.renderD(sc, 0)
    w = torch.linspace(-1.0, 1.0, img.numel(), device=img.device).view_as(img)
    (img * w).sum().backward()
    assert pose.grad is not None and pose.grad.shape == (4, 4) and torch.isfinite(pose.grad).all() and pose.grad.abs().max() > 0
    ctx = capi.Context(0)
    ctx.load_description(scene_io.load_scene_description(scene_path("cbox_bunny")), dict(width=40, height=40, spp=4, sppe=4, sppse=4))
    ctx.grad_require(capi.PARAM_SENSOR_TRANSFORM, 0)
    ctx.configure()
    ci = capi.make_integrator("direct", bsdf_samples=1, light_samples=1)
    assert torch.equal(img.detach(), ctx.render_d(ci))
    g = ctx.render_d_vjp(ci, w.contiguous()).view(4, 4)
    assert (pose.grad - g).norm() <= 1e-4 * g.norm()


def test_transform_gradient_contraction_matches_finite_differences(psdr_cuda):
    """psdr_cuda.Scene._transform_gradient (CPU, torch): dL/d(to_world_left|right) from an object-space vertex gradient must be
    the derivative of <g_world, world vertices> for world = L W R (x, 1) (mesh.cpp:223)."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(0)

    class FakeMesh:
        to_world_left = (np.eye(4) + 0.1 * rng.normal(size=(4, 4))).astype(np.float32)
        to_world_raw = (np.eye(4) + 0.1 * rng.normal(size=(4, 4))).astype(np.float32)
        to_world_right = (np.eye(4) + 0.1 * rng.normal(size=(4, 4))).astype(np.float32)
        vertex_positions = rng.normal(size=(7, 3)).astype(np.float32)
    for m in (FakeMesh.to_world_left, FakeMesh.to_world_raw, FakeMesh.to_world_right):
        m[3] = [0, 0, 0, 1]
    g_world = rng.normal(size=(7, 3))

    def world(L, R):
        M = L.astype(np.float64) @ FakeMesh.to_world_raw.astype(np.float64) @ R.astype(np.float64)
        x1 = np.concatenate([FakeMesh.vertex_positions.astype(np.float64), np.ones((7, 1))], axis=1)
        return (x1 @ M.T)[:, :3]

    M3 = (FakeMesh.to_world_left.astype(np.float64) @ FakeMesh.to_world_raw.astype(np.float64) @ FakeMesh.to_world_right.astype(np.float64))[:3, :3]
    g_obj = torch.tensor(g_world @ M3)                      # what the renderer returns: g_obj = M3^T g_world per vertex
    for left in (True, False):
        g = psdr_cuda.Scene._transform_gradient(torch, FakeMesh, left, g_obj).double().numpy()
        for (i, j) in ((0, 3), (1, 1), (2, 0), (0, 2)):
            h = 1e-4
            A = FakeMesh.to_world_left if left else FakeMesh.to_world_right
            Ap, Am = A.astype(np.float64).copy(), A.astype(np.float64).copy()
            Ap[i, j] += h; Am[i, j] -= h
            wp = world(Ap, FakeMesh.to_world_right) if left else world(FakeMesh.to_world_left, Ap)
            wm = world(Am, FakeMesh.to_world_right) if left else world(FakeMesh.to_world_left, Am)
            fd = float((g_world * (wp - wm)).sum() / (2 * h))
            assert abs(fd - g[i, j]) <= 1e-3 * max(1.0, abs(fd)), (left, i, j, fd, g[i, j])


def test_cpp_loader_on_a_uv_mapped_two_sensor_scene(psdr_cuda, textured_scene):
    """texture coordinates, uv face indices, a rotated mesh transform and a second sensor through the C++ ingest vs the oracle's loader
    (itself compared with the reference's loader + configure on this scene in tests/test_ref_render.py)"""
    from oracle import orc
    ref = orc.load_scene_description(textured_scene)
    sc = psdr_cuda.Scene(-1)
    sc.load_file(textured_scene, False)
    assert sc.num_sensors == 2 and sc.num_meshes == 3 and (sc.opts.width, sc.opts.height, sc.opts.spp) == (24, 16, 2)
    pm = sc.param_map
    for i, m in enumerate(ref["meshes"]):
        o = pm["Mesh[%d]" % i]
        assert np.array_equal(o.vertex_positions, m["verts"]) and np.array_equal(o.face_indices, m["faces"])
        assert np.allclose(o.to_world_raw, m["to_world"], atol=1e-7)
        if "uvs" in m:
            assert o.has_uv and np.array_equal(o.vertex_uv, m["uvs"]) and np.array_equal(o.face_uv_indices, m["uv_faces"])
        else:
            assert not o.has_uv
    for i, s in enumerate(ref["sensors"]):
        assert np.allclose(pm["Sensor[%d]" % i].to_world, s["to_world"], atol=1e-7) and pm["Sensor[%d]" % i].fov_x == np.float32(s["fov"])
    assert pm["Mesh[id=wall]"].bsdf.type_name() == "RoughConductor" and pm["BSDF[id=metal]"].alpha_u.data.reshape(-1)[0] == np.float32(0.3)


# ---- the module surface against the reference's own pybind11 module (src/psdr.cpp), and its utilities against the reference's own code ----
def test_python_surface_covers_the_reference_module(psdr_cuda):
    """tests/golden/ref_python_surface.json = every class / method / property psdr.cpp binds (tests/golden/make_ref_surface.py); each must
    exist here under the same name. The one entry this host does not offer is listed explicitly."""
    import json
    import os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "ref_python_surface.json")) as fh:
        surface = json.load(fh)
    ref_src = os.path.join(os.environ.get("PSDR_REFERENCE", "/root/reference"), "src", "psdr.cpp")
    if os.path.exists(ref_src):   # the fixture is what the reference binds today
        import importlib.util
        spec = importlib.util.spec_from_file_location("make_ref_surface", os.path.join(GOLDEN, "make_ref_surface.py"))
        gen = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(gen)
        with open(ref_src) as fh:
            assert gen.extract(fh.read()) == surface
    absent = set()                                       # every class, method and property src/psdr.cpp binds is present
    missing = []
    for cls, info in sorted(surface.items()):
        if not hasattr(psdr_cuda, cls):
            missing.append((cls, None))
            continue
        c = getattr(psdr_cuda, cls)
        try:
            target = c() if c.__module__.endswith("_surface") else c      # the numpy value types carry their fields on the instance
        except TypeError:
            target = c
        for name in info["methods"] + info["properties"]:
            if not hasattr(target, name) and (cls, name) not in absent:
                missing.append((cls, name))
    assert not missing, missing
    assert len(surface) == 28
    # inheritance the reference declares (py::class_<Derived, Base>)
    assert issubclass(psdr_cuda.DiffuseBSDF, psdr_cuda.BSDF) and issubclass(psdr_cuda.RoughConductorBSDF, psdr_cuda.BSDF)
    assert issubclass(psdr_cuda.AreaLight, psdr_cuda.Emitter) and issubclass(psdr_cuda.EnvironmentMap, psdr_cuda.Emitter)
    assert issubclass(psdr_cuda.PerspectiveCamera, psdr_cuda.Sensor) and issubclass(psdr_cuda.DirectIntegrator, psdr_cuda.Integrator)
    assert issubclass(psdr_cuda.PositionSampleC, psdr_cuda.SampleRecordC)


@pytest.fixture(scope="module")
def refrun():
    from oracle import refrun as r
    if not r.available():
        pytest.skip("oracle/_ref/libref_render.so not built and /root/reference absent")
    r.lib()
    return r


@pytest.mark.parametrize("name", ["cbox_bunny", "cbox_bunny_rc", "bunny_env", "bunny_env_2", "tree"])
def test_param_map_keys_and_type_names_match_reference_loader(psdr_cuda, refrun, name):
    """Scene::m_param_map as the reference's own SceneLoader fills it (scene_loader.cpp:184-240,343-352): same keys, same type_name()"""
    import ctypes as C
    import os
    from conftest import ROOT
    L = refrun.lib()
    L.ref_param_map_keys.restype = C.c_char_p
    r = refrun.Scene(scene_path(name), os.path.join(ROOT, "tests"))
    ref = dict(l.split("\t", 1)[0].rsplit("=", 1) for l in L.ref_param_map_keys(r.h).decode().strip().split("\n"))   # key=type_name<TAB>to_string
    sc = psdr_cuda.Scene(-1)
    sc.load_file(scene_path(name), False)
    mine = {k: v.type_name() for k, v in sc._raw_param_map().items()}
    assert mine == ref
    assert (sc.opts.width, sc.opts.height, sc.opts.spp) == (r.opts["width"], r.opts["height"], r.opts["spp"])


def test_distributions_match_reference_code(psdr_cuda, refrun):
    """DiscreteDistribution (pmf.cpp) and HyperCubeDistribution (cube_distrb.cpp) of the module against the reference's classes"""
    import ctypes as C
    L = refrun.lib()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 1000):
        pmf = rng.uniform(0, 2, n).astype(np.float32)
        if n > 3:
            pmf[2] = 0
        u = np.concatenate([rng.uniform(0, 1, 500), [0.0, 0.5, 0.999999]]).astype(np.float32)
        d = psdr_cuda.DiscreteDistribution()
        d.init(pmf)
        for reuse in (0, 1):
            idx, pdf, uo = np.zeros(len(u), np.int32), np.zeros(len(u), np.float32), np.zeros(len(u), np.float32)
            assert L.ref_discrete_sample(P(pmf), n, P(u), len(u), reuse, P(idx), P(pdf), P(uo)) == 0
            mine_u = u.copy()
            i2, p2 = d.sample_reuse(mine_u) if reuse else d.sample(mine_u)
            assert np.array_equal(np.broadcast_to(i2, idx.shape), idx) and np.allclose(np.broadcast_to(p2, pdf.shape), pdf, rtol=1e-6)
            if reuse:
                assert np.allclose(mine_u, uo, atol=2e-6)
        assert np.isclose(d.sum, pmf.astype(np.float64).sum(), rtol=1e-5)
    for ndim, reso in ((2, [5, 3]), (3, [4, 2, 3])):
        ncell = int(np.prod(reso))
        mass = rng.uniform(0, 1, ncell).astype(np.float32)
        s = rng.uniform(0, 1, (300, ndim)).astype(np.float32)
        warped, ps, pe = np.zeros_like(s), np.zeros(300, np.float32), np.zeros(300, np.float32)
        cells = np.zeros((ncell, ndim), np.int32)
        r = np.array(reso, np.int32)
        assert L.ref_hypercube(ndim, P(r), P(mass), P(s), 300, P(warped), P(ps), P(pe), P(cells)) == 0
        h = (psdr_cuda.HyperCubeDistribution2f if ndim == 2 else psdr_cuda.HyperCubeDistribution3f)()
        h.set_resolution(reso)
        h.set_mass(mass)
        assert np.array_equal(h.cells, cells)
        assert np.allclose(h.pdf(s), pe, rtol=1e-6)
        mine = s.copy()
        assert np.allclose(h.sample_reuse(mine), ps, rtol=1e-6) and np.allclose(mine, warped, atol=2e-6)


def test_bitmap_eval_matches_reference_code(psdr_cuda, refrun):
    """Bitmap1fD / Bitmap3fD .eval (bitmap.cpp:56-96) incl. the flipped v, wrap-around and the clamped last texel"""
    import ctypes as C
    L = refrun.lib()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(1)
    sc = psdr_cuda.Scene(-1)
    sc.load_file(scene_path("cbox_bunny_rc"), False)
    pm = sc._raw_param_map()
    uv = np.concatenate([rng.uniform(-1.5, 2.5, (400, 2)), [[0, 0], [1, 1], [0.999999, 0.5], [0.5, -1.0]]]).astype(np.float32)
    for bm, ch in ((pm["BSDF[0]"].reflectance, 3), (pm["BSDF[3]"].alpha_u, 1)):
        for (w, h) in ((1, 1), (7, 5), (2, 2)):
            tex = rng.uniform(0, 1, (h, w, ch)).astype(np.float32)
            bm.resolution = (w, h)
            bm.data = tex.reshape(-1, ch) if ch == 3 else tex.reshape(-1)
            for flip in (True, False):
                out = np.zeros((len(uv), ch), np.float32)
                assert L.ref_bitmap_eval(ch, w, h, P(tex), P(uv), len(uv), int(flip), P(out)) == 0, L.ref_last_error()
                mine = bm.eval(uv, flip)
                assert np.allclose(mine.reshape(len(uv), ch), out, atol=2e-6), (w, h, ch, flip)


def test_mesh_helpers_match_reference_code(psdr_cuda, refrun):
    """Mesh.vertex_normals, Mesh.edge_indices(), Mesh.sample_position and Mesh.bsdf against Mesh::configure / load / sample_position of
    the reference (mesh.cpp:19-51,143-203,277-303)"""
    import ctypes as C
    import os
    from conftest import ROOT
    L = refrun.lib()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    r = refrun.Scene(scene_path("cbox_bunny"), os.path.join(ROOT, "tests"), 8, 8, 1, 1, 1)
    r.configure()
    sc = psdr_cuda.Scene(-1)
    sc.load_file(scene_path("cbox_bunny"), False)
    pm = sc.param_map
    for i in (0, 1, 2):
        m = pm["Mesh[%d]" % i]
        vn = np.zeros((m.num_vertices, 3), np.float32)
        assert L.ref_mesh_vertex_normals(r.h, i, P(vn)) == 0
        assert np.abs(m.vertex_normals - vn).max() <= 5e-6
        e = r.mesh_edges(i)
        assert np.array_equal(m.edge_indices().T, e[:, :4])
        m.configure()
    assert pm["Mesh[1]"].bsdf.id == "white" and pm["Mesh[0]"].bsdf.type_name() == "Diffuse"
    s2 = np.random.default_rng(2).uniform(0, 1, (200, 2)).astype(np.float32)
    p, n, pdf = np.zeros((200, 3), np.float32), np.zeros((200, 3), np.float32), np.zeros(200, np.float32)
    assert L.ref_mesh_sample_position(r.h, 0, P(s2), 200, P(p), P(n), P(pdf)) == 0, L.ref_last_error()   # the emitter quad
    ps = pm["Mesh[0]"].sample_position(s2)
    assert np.abs(ps.p - p).max() <= 1e-4 and np.abs(ps.n - n).max() <= 1e-6 and np.allclose(ps.pdf, pdf, rtol=1e-6)
    assert isinstance(ps, psdr_cuda.PositionSampleC) and ps.is_valid.all() and np.all(ps.J == 1)
    f = psdr_cuda.FrameC(n)
    assert np.allclose(f.to_local(f.to_world(p)), p, atol=1e-3)
    ray = psdr_cuda.RayC(p, n)
    assert np.array_equal(ray.reversed().d, -n)


# ---- loader semantics on awkward XML, against the reference's own SceneLoader -----------------------------------------------------------------
def _describe_product(psdr_cuda, sc):
    """the same text oracle/ref_render_shim.cpp:ref_scene_describe prints for the reference's Scene, from this module's objects"""
    pm = sc._raw_param_map()
    fmt = lambda v: " %.9g" % float(v)
    mat = lambda m: "".join(fmt(x) for x in np.asarray(m, np.float32).reshape(-1))

    def tex(name, bm, ch):
        w, h = bm.resolution
        s = " %s[%dx%d]" % (name, w, h)
        if (w, h) == (1, 1):
            s += "".join(fmt(x) for x in np.asarray(bm.data, np.float32).reshape(-1)[:ch])
        return s
    out = ["opts %d %d %d %d %d" % (sc.opts.width, sc.opts.height, sc.opts.spp, sc.opts.sppe, sc.opts.sppse)]
    for i in range(sc.num_sensors):
        s = pm["Sensor[%d]" % i]
        out.append("sensor" + fmt(s.fov_x) + fmt(s.near_clip) + fmt(s.far_clip) + mat(s.to_world))
    i = 0
    while "BSDF[%d]" % i in pm:
        b = pm["BSDF[%d]" % i]
        line = "bsdf %s id=%s" % (b.type_name(), b.id)
        if b.type_name() == "Diffuse":
            line += tex("reflectance", b.reflectance, 3)
        else:
            line += tex("alpha_u", b.alpha_u, 1) + tex("alpha_v", b.alpha_v, 1) + tex("eta", b.eta, 3) + tex("k", b.k, 3) + tex("specular_reflectance", b.specular_reflectance, 3)
        out.append(line)
        i += 1
    emitters = []
    i = 0
    while "Emitter[%d]" % i in pm:
        emitters.append(pm["Emitter[%d]" % i])
        i += 1
    for e in emitters:
        if isinstance(e, psdr_cuda.EnvironmentMap):
            w, h = e.radiance.resolution
            out.append("envmap" + fmt(e.scale) + " %dx%d" % (w, h) + mat(e.to_world))
    i = 0
    while "Mesh[%d]" % i in pm:
        m = pm["Mesh[%d]" % i]
        line = "mesh id=%s nv=%d nf=%d uv=%d face_normals=%d edges=%d bsdf=%s" % (m.id, m.num_vertices, m.num_faces, len(m.vertex_uv) if m.has_uv else 0, int(m.use_face_normals),
                                                                                  int(m.enable_edges), pm["BSDF[%d]" % m.bsdf_index].id if m.bsdf_index >= 0 else "-")
        if m.emitter_index >= 0:
            line += " radiance" + "".join(fmt(x) for x in emitters[m.emitter_index].radiance)
        out.append(line + " to_world" + mat(m.to_world_raw))
        i += 1
    return "\n".join(out) + "\n"


_SENSOR = """<sensor type="perspective"><float name="fov" value="%s"/>%s
  <transform name="%s">%s</transform>
  <sampler type="independent"><integer name="sampleCount" value="3"/></sampler>
  <film type="hdrfilm"><integer name="width" value="20"/><integer name="height" value="10"/></film></sensor>"""
_SHAPE = """<shape type="obj" id="%s"><string name="filename" value="./data/objects/cbox/%s.obj"/>%s<ref id="%s"/>%s</shape>"""
LOADER_CASES = {
    # aliases (toWorld / lookAt / fovAxis / nearClip / farClip / faceNormals), attribute defaults of translate / scale, float and short rgb values
    "aliases": "<scene>" + _SENSOR % ("35.5", '<string name="fovAxis" value="x"/><float name="nearClip" value="0.5"/><float name="farClip" value="250"/>', "toWorld",
                                      '<lookAt origin="1, 2, 3" target="0, 0.5, 0" up="0, 1, 0"/>')
    + '<bsdf type="diffuse" id="a"><float name="reflectance" value="0.3"/></bsdf><bsdf type="diffuse" id="b"><rgb name="reflectance" value="0.7"/></bsdf>'
    + '<bsdf type="diffuse" id="c"><rgb name="reflectance" value="0.1, 0.6"/></bsdf>'
    + _SHAPE % ("f", "floor", '<transform name="toWorld"><scale x="2"/><translate y="-1.5"/><rotate x="0" y="0" z="1" angle="33"/></transform>', "a", "")
    + _SHAPE % ("l", "emitter", '<boolean name="faceNormals" value="true"/>', "c", '<emitter type="area"><rgb name="radiance" value="5"/></emitter>') + "</scene>",
    # matrix transforms, a transform sequence applied left to right, lookat spelled lowercase, a second sensor without film / sampler
    "matrix": "<scene>" + _SENSOR % ("60", "", "to_world", '<matrix value="1 0 0 0.5  0 1 0 -2  0 0 1 7  0 0 0 1"/><translate x="1" z="-1"/>')
    + '<sensor type="perspective"><float name="fov" value="20"/><transform name="to_world"><lookat origin="0, 1, 9" target="0, 1, 0" up="0, 1, 0"/></transform></sensor>'
    + '<bsdf type="roughconductor" id="m"><float name="alpha" value="0.25"/><rgb name="eta" value="0.2, 0.9, 1.1"/><rgb name="k" value="3.9"/></bsdf>'
    + _SHAPE % ("w", "wall_back", '<transform name="to_world"><rotate x="1" y="0" z="0" angle="-90"/><scale x="0.5" y="2" z="3"/><matrix value="0 1 0 0  1 0 0 0  0 0 1 0  0 0 0 1"/></transform>', "m", "")
    + _SHAPE % ("l", "emitter", "", "m", '<emitter type="area"><rgb name="radiance" value="1, 2, 3"/></emitter>') + "</scene>",
}
# bitmap textures (texture type="bitmap" + filename) on a diffuse reflectance and a rough-conductor roughness, and an environment map with a
# scale and a transform
LOADER_CASES["textures"] = ("<scene>" + _SENSOR % ("45", "", "to_world", '<translate z="5"/>')
    + '<bsdf type="diffuse" id="t"><texture name="reflectance" type="bitmap"><string name="filename" value="./data/envmaps/ballroom_1k.exr"/></texture></bsdf>'
    + '<bsdf type="roughconductor" id="r"><texture name="alpha" type="bitmap"><string name="filename" value="./data/envmaps/ballroom_1k.exr"/></texture>'
    + '<rgb name="eta" value="0.2, 0.9, 1.1"/><rgb name="k" value="3.9, 2.4, 2.2"/></bsdf>'
    + '<emitter type="envmap"><string name="filename" value="./data/envmaps/ballroom_1k.exr"/><float name="scale" value="2.5"/>'
    + '<transform name="to_world"><rotate x="0" y="1" z="0" angle="90"/></transform></emitter>'
    + _SHAPE % ("f", "floor", "", "t", "") + _SHAPE % ("w", "wall_back", "", "r", "") + "</scene>")
LOADER_ERRORS = {
    "unknown bsdf ref": "<scene>" + _SENSOR % ("30", "", "to_world", "") + '<bsdf type="diffuse" id="a"><float name="reflectance" value="0.3"/></bsdf>' + _SHAPE % ("f", "floor", "", "zzz", "") + "</scene>",
    "duplicate bsdf id": "<scene>" + _SENSOR % ("30", "", "to_world", "") + '<bsdf type="diffuse" id="a"><float name="reflectance" value="0.3"/></bsdf><bsdf type="diffuse" id="a"><float name="reflectance" value="0.4"/></bsdf>' + "</scene>",
    "fov axis y": "<scene>" + _SENSOR % ("30", '<string name="fov_axis" value="y"/>', "to_world", "") + "</scene>",
    "unsupported bsdf": "<scene>" + _SENSOR % ("30", "", "to_world", "") + '<bsdf type="plastic" id="a"/>' + "</scene>",
    "unsupported transform": "<scene>" + _SENSOR % ("30", "", "to_world", '<shear x="1"/>') + "</scene>",
    "bad transform name": "<scene>" + _SENSOR % ("30", "", "world", "") + "</scene>",
    "second film": "<scene>" + _SENSOR % ("30", "", "to_world", "") + _SENSOR % ("30", "", "to_world", "") + "</scene>",
    "missing reflectance": "<scene>" + _SENSOR % ("30", "", "to_world", "") + '<bsdf type="diffuse" id="a"/>' + "</scene>",
    "short vector": "<scene>" + _SENSOR % ("30", "", "to_world", '<lookat origin="1, 2" target="0, 0, 0" up="0, 1, 0"/>') + "</scene>",
}


@pytest.mark.parametrize("case", sorted(LOADER_CASES))
def test_loader_semantics_match_reference_loader(psdr_cuda, refrun, monkeypatch, case):
    """scene_loader.cpp:24-419 on XML the fixtures do not contain: what the reference's own SceneLoader::load_from_string makes of it vs
    this module's C++ ingest, object by object and number by number"""
    import ctypes as C
    import os
    from conftest import ROOT
    L = refrun.lib()
    L.ref_scene_describe.restype = C.c_char_p
    L.ref_scene_load_string.restype = C.c_void_p
    L.ref_scene_load_string.argtypes = [C.c_char_p, C.c_char_p]
    tests_dir = os.path.join(ROOT, "tests")
    h = L.ref_scene_load_string(LOADER_CASES[case].encode(), tests_dir.encode())
    assert h, L.ref_last_error()
    ref = L.ref_scene_describe(C.c_void_p(h)).decode()
    L.ref_scene_free(C.c_void_p(h))
    monkeypatch.chdir(tests_dir)
    sc = psdr_cuda.Scene(-1)
    sc.load_string(LOADER_CASES[case], False)
    mine = _describe_product(psdr_cuda, sc)
    rl, ml = ref.strip().split("\n"), mine.strip().split("\n")
    assert len(rl) == len(ml), (ref, mine)
    for a, b in zip(rl, ml):
        ta, tb = a.split(), b.split()
        assert len(ta) == len(tb), (a, b)
        for x, y in zip(ta, tb):
            try:
                fx, fy = float(x), float(y)
            except ValueError:
                assert x == y, (a, b)
                continue
            assert abs(fx - fy) <= 2e-6 * max(1.0, abs(fx)), (a, b)


@pytest.mark.parametrize("case", sorted(LOADER_ERRORS))
def test_loader_errors_match_reference_loader(psdr_cuda, refrun, monkeypatch, case):
    """malformed scenes the reference's loader rejects (PSDR_ASSERT_MSG in scene_loader.cpp) are rejected here too"""
    import ctypes as C
    import os
    from conftest import ROOT
    L = refrun.lib()
    L.ref_scene_load_string.restype = C.c_void_p
    L.ref_scene_load_string.argtypes = [C.c_char_p, C.c_char_p]
    tests_dir = os.path.join(ROOT, "tests")
    h = L.ref_scene_load_string(LOADER_ERRORS[case].encode(), tests_dir.encode())
    assert not h, "the reference accepted: " + case
    monkeypatch.chdir(tests_dir)
    with pytest.raises(RuntimeError):
        psdr_cuda.Scene(-1).load_string(LOADER_ERRORS[case], False)


@pytest.mark.parametrize("case", sorted(LOADER_CASES))
def test_python_loaders_agree_on_the_awkward_xml(monkeypatch, case):
    """the ctypes path's loader (scene_io.py) and the checker's (orc.py) read the same XML to the same description"""
    import os
    from conftest import ROOT
    from oracle import orc
    from psdr_cuda_b200 import scene_io
    monkeypatch.chdir(os.path.join(ROOT, "tests"))
    a, b = scene_io.load_scene_description(xml_string=LOADER_CASES[case]), orc.load_scene_description(xml_string=LOADER_CASES[case])
    assert a["opts"] == b["opts"] and len(a["meshes"]) == len(b["meshes"]) and len(a["sensors"]) == len(b["sensors"])
    for x, y in zip(a["sensors"], b["sensors"]):
        assert np.array_equal(x["to_world"], y["to_world"]) and (x["fov"], x["near"], x["far"]) == (y["fov"], y["near"], y["far"])
    for x, y in zip(a["meshes"], b["meshes"]):
        assert np.array_equal(x["to_world"], y["to_world"]) and np.array_equal(x["verts"], y["verts"]) and x["bsdf"] == y["bsdf"] and x["face_normals"] == y["face_normals"]
    for x, y in zip(a["bsdfs"], b["bsdfs"]):
        assert x["type"] == y["type"] and all(np.array_equal(x[k], y[k]) for k in x if isinstance(x[k], np.ndarray))


@pytest.mark.parametrize("case", sorted(LOADER_ERRORS))
def test_python_loaders_reject_what_the_reference_rejects(monkeypatch, case):
    import os
    from conftest import ROOT
    from oracle import orc
    from psdr_cuda_b200 import scene_io
    monkeypatch.chdir(os.path.join(ROOT, "tests"))
    for load in (scene_io.load_scene_description, orc.load_scene_description):
        with pytest.raises(Exception):
            load(xml_string=LOADER_ERRORS[case])


def test_mesh_dump_writes_the_reference_file(psdr_cuda, refrun, textured_scene, tmp_path):
    """Mesh::dump (mesh.cpp:318-392): smooth-normal, face-normal and uv-mapped meshes come out as the files the reference writes (numbers
    in %.6e; a vertex normal may differ in its last printed digit)"""
    import os
    from conftest import ROOT
    L = refrun.lib()
    for xml, meshes in ((scene_path("cbox_bunny"), (0, 1)), (textured_scene, (0, 2))):
        r = refrun.Scene(xml, os.path.join(ROOT, "tests"), 8, 8, 1, 0, 0)
        r.configure()
        sc = psdr_cuda.Scene(-1)
        sc.load_file(xml, False)
        for i in meshes:
            a, b = str(tmp_path / "ref.obj"), str(tmp_path / "mine.obj")
            assert L.ref_mesh_dump(r.h, i, a.encode()) == 0, L.ref_last_error()
            sc.param_map["Mesh[%d]" % i].dump(b)
            la, lb = open(a).read().split("\n"), open(b).read().split("\n")
            assert len(la) == len(lb)
            for x, y in zip(la, lb):
                if x != y:
                    tx, ty = x.split(), y.split()
                    assert tx[0] == ty[0] == "vn" and np.allclose([float(v) for v in tx[1:]], [float(v) for v in ty[1:]], atol=2e-6), (x, y)


def test_reprs_match_reference_to_string(psdr_cuda, refrun):
    """Object::to_string of the reference's scene objects (what repr() shows in its Python module) for BSDFs, meshes, sensors and the
    environment map; AreaLight prints an Enoki array and is left out"""
    import ctypes as C
    import os
    from conftest import ROOT
    L = refrun.lib()
    L.ref_param_map_keys.restype = C.c_char_p
    for name in ("cbox_bunny_rc", "bunny_env", "tree"):
        r = refrun.Scene(scene_path(name), os.path.join(ROOT, "tests"))
        sc = psdr_cuda.Scene(-1)
        sc.load_file(scene_path(name), False)
        pm = sc.param_map
        n = 0
        for line in L.ref_param_map_keys(r.h).decode().strip().split("\n"):
            head, rep = line.split("\t", 1)
            key = head[:head.index("]=") + 1]
            if rep.startswith("AreaLight["):
                continue
            assert repr(pm[key]) == rep, (key, rep, repr(pm[key]))
            n += 1
        assert n >= 4
    assert repr(psdr_cuda.RenderOption(3, 4, 5)) == "[width: 3, height: 4, spp: 5, sppe: 5, sppse: 5, log_level: 1]"


def test_face_indices_are_writable_before_the_first_configure(psdr_cuda):
    """src/psdr.cpp:255-256: face_indices / face_uv_indices are read-write; here until the mesh has gone to the device"""
    sc = psdr_cuda.Scene(-1)
    sc.load_file(scene_path("cbox_bunny"), False)
    m = sc.param_map["Mesh[2]"]
    f = m.face_indices.copy()
    m.face_indices = f[:, [1, 2, 0]]
    assert np.array_equal(m.face_indices, f[:, [1, 2, 0]])
    assert np.abs(m.vertex_normals - sc.param_map["Mesh[2]"].vertex_normals).max() == 0
    with pytest.raises(RuntimeError, match="out of range"):
        m.face_indices = f + 100
    with pytest.raises(RuntimeError):
        m.face_indices = f.reshape(-1)


def test_bitmap_constructors_and_writable_members(psdr_cuda):
    """Bitmap1fD / Bitmap3fD (), (value), (width, height, data), (file name) (src/psdr.cpp:102-106,112-116) and the read-write bitmap members
    of the BSDFs and the environment map (src/psdr.cpp:208-215,236)"""
    import os
    from conftest import DATA
    assert psdr_cuda.Bitmap3fD().resolution == (1, 1) and psdr_cuda.Bitmap1fD().channels == 1
    assert np.allclose(psdr_cuda.Bitmap3fD([0.1, 0.2, 0.3]).data, [[0.1, 0.2, 0.3]]) and np.allclose(psdr_cuda.Bitmap1fD(0.25).data, [[0.25]])
    t = psdr_cuda.Bitmap3fD(2, 3, np.arange(18, dtype=np.float32).reshape(6, 3))
    assert t.resolution == (2, 3) and np.allclose(t.eval(np.array([[0.5, 0.5]], np.float32)), [[7.5, 8.5, 9.5]])
    assert psdr_cuda.Bitmap3fD(os.path.join(DATA, "envmaps", "ballroom_1k.exr")).resolution == (1024, 512)
    sc = psdr_cuda.Scene(-1)
    sc.load_file(scene_path("cbox_bunny_rc"), False)
    pm = sc.param_map
    pm["BSDF[0]"].reflectance = t
    assert pm["BSDF[0]"].reflectance.resolution == (2, 3) and np.array_equal(pm["BSDF[0]"].reflectance.data, t.data)
    pm["BSDF[3]"].alpha_u = psdr_cuda.Bitmap1fD(0.4)
    assert np.allclose(pm["BSDF[3]"].alpha_u.data, 0.4)
    with pytest.raises(RuntimeError, match="channel"):
        pm["BSDF[3]"].alpha_u = t
    sc2 = psdr_cuda.Scene(-1)
    sc2.load_file(scene_path("bunny_env"), False)
    sc2.param_map["Emitter[0]"].radiance = psdr_cuda.Bitmap3fD(4, 2, np.ones((8, 3), np.float32))
    assert sc2.param_map["Emitter[0]"].radiance.resolution == (4, 2)


def test_every_class_the_reference_can_construct_from_python_can_be_constructed(psdr_cuda):
    """py::init<> overloads of src/psdr.cpp (the fixture counts them per class): each such class has a working constructor here"""
    import json
    import os
    from conftest import DATA, GOLDEN
    with open(os.path.join(GOLDEN, "ref_python_surface.json")) as fh:
        surface = json.load(fh)
    mesh = psdr_cuda.Mesh()
    mesh.load(os.path.join(DATA, "objects", "cbox", "emitter.obj"))
    exr = os.path.join(DATA, "envmaps", "ballroom_1k.exr")
    args = {"RenderOption": (32, 16, 4), "RayC": (), "RayD": (), "FrameC": (np.array([[0, 0, 1.0]], np.float32),), "FrameD": (), "Bitmap1fD": (0.5,), "Bitmap3fD": ([0.1, 0.2, 0.3],),
            "DiscreteDistribution": (), "HyperCubeDistribution2f": (), "HyperCubeDistribution3f": (), "PerspectiveCamera": (40.0, 0.1, 1e3), "AreaLight": ([1.0, 2.0, 3.0], mesh),
            "EnvironmentMap": (exr,), "Mesh": (), "Scene": (-1,), "FieldExtractionIntegrator": ("depth",), "DirectIntegrator": (2, 1)}
    constructible = sorted(c for c, info in surface.items() if info["constructors"] > 0)
    assert constructible == sorted(args)
    for cls in constructible:
        obj = getattr(psdr_cuda, cls)(*args[cls])
        assert obj is not None
    cam = psdr_cuda.PerspectiveCamera(40.0, 0.1, 1e3)
    assert cam.fov_x == 40.0 and np.array_equal(cam.to_world, np.eye(4, dtype=np.float32))
    assert psdr_cuda.EnvironmentMap(exr).radiance.resolution == (1024, 512) and psdr_cuda.AreaLight([1.0, 2.0, 3.0], mesh).radiance == (1.0, 2.0, 3.0)


def test_keyword_arguments_of_the_reference_bindings(psdr_cuda, monkeypatch):
    """the "name"_a keywords of src/psdr.cpp (RenderOption, Mesh.load / set_transform / append_transform / sample_position, Scene.load_file /
    load_string, Bitmap.eval, DirectIntegrator) are accepted under the same names and defaults"""
    import os
    from conftest import DATA, ROOT
    o = psdr_cuda.RenderOption(width=8, height=4, spp=2, sppe=1, sppse=3)
    assert (o.width, o.height, o.spp, o.sppe, o.sppse) == (8, 4, 2, 1, 3)
    assert psdr_cuda.RenderOption(width=8, height=4, spp=2, sppe=5).sppse == 5
    m = psdr_cuda.Mesh()
    m.load(filename=os.path.join(DATA, "objects", "cbox", "emitter.obj"), verbose=False)
    m.set_transform(mat=np.eye(4, dtype=np.float32), set_left=False)
    m.append_transform(mat=np.eye(4, dtype=np.float32), append_left=True)
    assert m.sample_position(sample2=np.array([[0.3, 0.6]], np.float32), active=True).p.shape == (1, 3)
    sc = psdr_cuda.Scene(-1)
    sc.load_file(file_name=scene_path("cbox_bunny"), auto_configure=False)
    monkeypatch.chdir(os.path.join(ROOT, "tests"))
    sc2 = psdr_cuda.Scene(-1)
    sc2.load_string(scene_xml=LOADER_CASES["aliases"], auto_configure=False)
    assert sc2.num_meshes == 2
    assert psdr_cuda.Bitmap3fD([0.5, 0.5, 0.5]).eval(uv=np.zeros((2, 2), np.float32), flip_v=False).shape == (2, 3)
    d = psdr_cuda.DirectIntegrator(bsdf_samples=2, light_samples=3)
    assert not d.hide_emitters
    import inspect
    assert list(inspect.signature(psdr_cuda.Integrator.renderC).parameters)[1:] == ["scene", "sensor_id"]
    assert list(inspect.signature(psdr_cuda.Integrator.renderD).parameters)[1:] == ["scene", "sensor_id"]
    assert "nrounds" in psdr_cuda.Integrator.preprocess_secondary_edges.__doc__ and "resolution" in psdr_cuda.Integrator.preprocess_secondary_edges.__doc__
