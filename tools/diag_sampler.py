"""GPU diagnostic: compare the sampler's internal state after k rounds with the oracle's trace."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hold_b200 import capi, scene_io, synth
from hold_b200.model import ErrorBoundSampler
from oracle import hold_oracle as O

S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
beta = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
add_tiny = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-6
ctx = capi.Context(0)
dev = torch.device("cuda", 0)
L = capi.lib()
L.hold_debug_ws_copy.restype = C.c_int
L.hold_debug_ws_copy.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]


def ws(slot, shape):
    t = torch.empty(shape, device=dev)
    rc = L.hold_debug_ws_copy(ctx.h, slot, C.c_void_p(t.data_ptr()), C.c_size_t(t.numel() * 4))
    assert rc == 0, rc
    return t.cpu()


for nid_sel in ("right", "object"):
  for k in range(1, 6):
    sc = synth.make_scene(H=12, W=12, S=S, nodes=("right", "object"), seed=3)
    for nid in sc.node_ids:
        sc.beta[nid] = torch.tensor(beta)
    sc.sampler["max_total_iters"] = k
    sc.sampler["add_tiny"] = add_tiny
    net = scene_io.build_net(sc, ctx)
    inp = scene_io.scene_input(sc, dev)
    art = O.scene_articulation(sc)
    a = art[nid_sel]
    dirs, cam = O.camera_rays(sc.uv, sc.extrinsics, sc.intrinsics)
    P = dirs.shape[1]
    dirs = dirs.reshape(-1, 3); cam = cam.unsqueeze(1).repeat(1, P, 1).reshape(-1, 3)
    frame = torch.zeros(P, dtype=torch.long)
    tr = []
    f = O.node_forward(a["kind"], O.CLASS_ID[nid_sel], dirs, cam, frame, sc.sdf_state[nid_sel], sc.rgb_state[nid_sel], sc.beta[nid_sel],
                       sc.sampler, sc.bounding_sphere, a["tfs"], posed_verts=a.get("verts") if a["kind"] == "hand" else None,
                       cano_verts=a.get("cano_verts"), skin_W=a.get("skin_W"), pose_cond=a.get("pose_cond"),
                       time_code=sc.time_code if a["kind"] == "object" else None, trace=tr)
    node = net.nodes[nid_sel]
    pose, keep, _, _ = node.articulate(inp)
    z, iters = ErrorBoundSampler(node).get_z_vals(dirs.to(dev), cam.to(dev), pose, 1)
    ctx.check()
    it_g = int(iters.item())
    last = tr[-1]
    n = last["z"].shape[1]
    R = dirs.shape[0]
    Ne = sc.sampler["N_samples_eval"]
    if it_g == f["iters"] == k or True:
        zg = ws(0, (R, 640))[:, :n]; sg = ws(1, (R, 640))[:, :n]; bg = ws(4, (R,))
        dz = (zg - last["z"]).abs(); ds = (sg - last["sdf"]).abs(); db = (bg - last["beta"]).abs()
        dzo = (z.cpu() - f["z_vals"]).abs()
        print(f"{nid_sel} k={k} iters gpu {it_g} oracle {f['iters']} n={n} | z max {dz.max():.2e} frac>1e-5 {(dz>1e-5).float().mean():.4f} | sdf max {ds.max():.2e} frac>1e-5 {(ds>1e-5).float().mean():.4f} | beta max {db.max():.2e} rel {(db/last['beta']).max():.2e} | zout frac>8e-4 {(dzo>8e-4).float().mean():.4f} max {dzo.max():.2e}")
