import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hold_b200 import capi, scene_io, synth
from hold_b200.model import ErrorBoundSampler
from oracle import hold_oracle as O
torch.set_printoptions(precision=5, linewidth=200, sci_mode=False)
S, beta, k, nid_sel = 32, 0.1, 3, "object"
ctx = capi.Context(0); dev = torch.device("cuda", 0); L = capi.lib()
L.hold_debug_ws_copy.restype = C.c_int
L.hold_debug_ws_copy.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
def ws(slot, shape):
    t = torch.empty(shape, device=dev)
    assert L.hold_debug_ws_copy(ctx.h, slot, C.c_void_p(t.data_ptr()), C.c_size_t(t.numel() * 4)) == 0
    return t.cpu()
sc = synth.make_scene(H=12, W=12, S=S, nodes=("right", "object"), seed=3)
for nid in sc.node_ids: sc.beta[nid] = torch.tensor(beta)
sc.sampler["max_total_iters"] = k; sc.sampler["add_tiny"] = 1e-3
net = scene_io.build_net(sc, ctx); inp = scene_io.scene_input(sc, dev)
art = O.scene_articulation(sc); a = art[nid_sel]
dirs, cam = O.camera_rays(sc.uv, sc.extrinsics, sc.intrinsics); P = dirs.shape[1]
dirs = dirs.reshape(-1, 3).contiguous(); cam = cam.unsqueeze(1).repeat(1, P, 1).reshape(-1, 3).contiguous()
tr = []
f = O.node_forward(a["kind"], O.CLASS_ID[nid_sel], dirs, cam, torch.zeros(P, dtype=torch.long), sc.sdf_state[nid_sel], sc.rgb_state[nid_sel], sc.beta[nid_sel],
                   sc.sampler, sc.bounding_sphere, a["tfs"], time_code=sc.time_code, trace=tr)
node = net.nodes[nid_sel]
pose, keep, _, _ = node.articulate(inp)
z, iters = ErrorBoundSampler(node).get_z_vals(dirs.to(dev), cam.to(dev), pose, 1)
ctx.check()
R = dirs.shape[0]
zn = ws(2, (R, 32))
so = tr[1]["samples"]
d = (torch.sort(zn, 1).values - torch.sort(so, 1).values).abs()
print("round-2 samples: frac>1e-5", (d > 1e-5).float().mean().item(), "max", d.max().item())
r = d.max(1).values.argmax().item()
print("worst ray", r)
print("gpu   ", zn[r])
print("oracle", so[r])
print("oracle inds", tr[1]["inds"][r])
print("oracle z(64)", tr[1]["z"][r])
print("oracle beta", tr[1]["beta"][r].item(), "gpu beta (after round3)", ws(4, (R,))[r].item())
zb = ws(0, (R, 640))[r, :96]
print("gpu zbuf(96)", zb)
print("oracle z3(96)", tr[2]["z"][r])
