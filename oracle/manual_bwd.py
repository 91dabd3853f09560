"""Closed-form reverse mode of the two pose servers — TEST INFRASTRUCTURE, like everything under oracle/.

The reference differentiates `GenericServer.forward` / `ObjectModel.forward` with torch autograd
(`fitting/model.py:117`, `optimize_ckpt.py`): the oracle for a gradient is therefore `torch.autograd` over
`oracle.hold_oracle.mano_server` / `object_server`.  This file restates the same gradients in closed form, step
by step in the order the CUDA kernels (`hold_b200/csrc/pose_bwd.cuh`) evaluate them, so that the derivation the
kernels implement is itself checked against autograd on the CPU (tests/test_cpu_pose_bwd.py)."""
import torch

from . import hold_oracle as O


def rodrigues_bwd(rv, gR):
    """d/d rv of `batch_rodrigues` (utils/external/lbs.py:298-329).  rv [3], gR [3,3] -> [3]."""
    a = rv + 1e-8
    n = torch.sqrt((a * a).sum())
    u = rv / n
    s, c = torch.sin(n), torch.cos(n)
    K = torch.tensor([[0.0, -u[2], u[1]], [u[2], 0.0, -u[0]], [-u[1], u[0], 0.0]], dtype=rv.dtype)
    KK = K @ K
    g_n = (gR * (c * K + s * KK)).sum()
    M = s * gR + (1 - c) * (gR @ K.T + K.T @ gR)
    g_u = torch.stack([M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1]])
    return g_n * a / n + g_u / n - a * (rv * g_u).sum() / n**3


def mano_server_bwd(m, scene_scale, transl, full_pose, betas, tfs_c_inv, g_verts=None, g_jnts=None, g_tfs=None):
    """Reverse mode of oracle.mano_server.  Returns g_betas [B,10], g_pose [B,48], g_transl [B,3], g_scale [B]."""
    B = full_pose.shape[0]
    dt = full_pose.dtype
    par = [int(p) for p in m["parents"]]
    Sd = m["shapedirs"].reshape(-1, 10).to(dt)        # [2334,10]
    Pd = m["posedirs"].to(dt)                         # [135,2334]
    Jr = m["J_regressor"].to(dt)                      # [16,778]
    W = m["lbs_weights"].to(dt)                       # [778,16]
    out = [torch.zeros(B, 10, dtype=dt), torch.zeros(B, 48, dtype=dt), torch.zeros(B, 3, dtype=dt), torch.zeros(B, dtype=dt)]
    for b in range(B):
        # ---- forward intermediates (same order as k_mano_lbs)
        pose = full_pose[b] + torch.cat([torch.zeros(3, dtype=dt), m["hands_mean"].to(dt)])
        vs = m["v_template"].to(dt) + (Sd @ betas[b]).view(-1, 3)
        J = Jr @ vs
        R = O.rodrigues(pose.view(16, 3)).to(dt)
        pf = (R[1:] - torch.eye(3, dtype=dt)).reshape(-1)
        vp = vs + (pf @ Pd).view(-1, 3)
        G = []
        for i in range(16):
            T = torch.eye(4, dtype=dt)
            T[:3, :3] = R[i]
            T[:3, 3] = J[i] - (J[par[i]] if i > 0 else 0)
            G.append(T if i == 0 else G[par[i]] @ T)
        G = torch.stack(G)
        A = G.clone()
        A[:, :3, 3] = G[:, :3, 3] - torch.einsum("irc,ic->ir", G[:, :3, :3], J)
        Tv = torch.einsum("vj,jrc->vrc", W, A[:, :3, :])           # [778,3,4]
        o = torch.einsum("vrc,vc->vr", Tv[:, :, :3], vp) + Tv[:, :, 3]
        s, t = scene_scale[b], transl[b]
        # ---- reverse
        gv = torch.zeros(778, 3, dtype=dt) if g_verts is None else g_verts[b].clone()
        gGt = torch.zeros(16, 3, dtype=dt)
        g_s = torch.zeros((), dtype=dt)
        g_t = torch.zeros(3, dtype=dt)
        if g_jnts is not None:
            for k, tip in enumerate(m["tip_ids"].tolist()):
                gv[tip] += g_jnts[b, 16 + k]
            gj = g_jnts[b, :16]
            gGt += s * gj
            g_t += s * gj.sum(0)
            g_s += (gj * (G[:, :3, 3] + t)).sum()
        g_o = s * gv
        g_t += s * gv.sum(0)
        g_s += (gv * (o + t)).sum()
        gA = torch.zeros(16, 3, 4, dtype=dt)
        if g_tfs is not None:
            gAs = g_tfs[b] if tfs_c_inv is None else torch.einsum("nik,njk->nij", g_tfs[b], tfs_c_inv.to(dt))
            gA += s * gAs[:, :3, :]
            g_s += (gAs[:, :3, :] * A[:, :3, :]).sum() + (gAs[:, :3, 3] * t).sum()
            g_t += s * gAs[:, :3, 3].sum(0)
        # skinning
        gA[:, :, :3] += torch.einsum("vj,vr,vc->jrc", W, g_o, vp)
        gA[:, :, 3] += torch.einsum("vj,vr->jr", W, g_o)
        g_vp = torch.einsum("vrc,vr->vc", Tv[:, :, :3], g_o)
        # A -> G, J
        gGR = gA[:, :, :3] - torch.einsum("ir,ic->irc", gA[:, :, 3], J)
        gGt = gGt + gA[:, :, 3]
        gJ = -torch.einsum("irc,ir->ic", G[:, :3, :3], gA[:, :, 3])
        # kinematic chain, children before parents
        gR = torch.zeros(16, 3, 3, dtype=dt)
        for i in range(15, -1, -1):
            if i == 0:
                gTR, gTt = gGR[0], gGt[0]
            else:
                p = par[i]
                GpR = G[p, :3, :3]
                TR, Tt = R[i], J[i] - J[p]
                gTR, gTt = GpR.T @ gGR[i], GpR.T @ gGt[i]
                gGR[p] = gGR[p] + gGR[i] @ TR.T + torch.outer(gGt[i], Tt)
                gGt[p] = gGt[p] + gGt[i]
                gJ[p] = gJ[p] - gTt
            gR[i] = gR[i] + gTR
            gJ[i] = gJ[i] + gTt
        # pose blend shapes
        g_vs = g_vp.clone()
        g_pf = Pd @ g_vp.reshape(-1)
        gR[1:] = gR[1:] + g_pf.view(15, 3, 3)
        # joint regressor, shape blend shapes
        g_vs = g_vs + Jr.T @ gJ
        out[0][b] = Sd.T @ g_vs.reshape(-1)
        out[1][b] = torch.cat([rodrigues_bwd(pose[3 * j: 3 * j + 3], gR[j]) for j in range(16)])
        out[2][b] = g_t
        out[3][b] = g_s
    return tuple(out)


def axis_angle_to_matrix_bwd(aa, gR):
    """d/d aa of common/rot.py axis_angle_to_matrix (quaternion route).  aa [3], gR [3,3] -> [3]."""
    dt = aa.dtype
    ang = torch.sqrt((aa * aa).sum())
    h = 0.5 * ang
    small = bool(ang.abs() < 1e-6)
    k = (0.5 - ang * ang / 48) if small else torch.sin(h) / ang
    qr, qi, qj, qk = torch.cos(h), aa[0] * k, aa[1] * k, aa[2] * k
    N = qr * qr + qi * qi + qj * qj + qk * qk
    ts = 2.0 / N
    g = gR
    g_ts = (-g[0, 0] * (qj * qj + qk * qk) + g[0, 1] * (qi * qj - qk * qr) + g[0, 2] * (qi * qk + qj * qr)
            + g[1, 0] * (qi * qj + qk * qr) - g[1, 1] * (qi * qi + qk * qk) + g[1, 2] * (qj * qk - qi * qr)
            + g[2, 0] * (qi * qk - qj * qr) + g[2, 1] * (qj * qk + qi * qr) - g[2, 2] * (qi * qi + qj * qj))
    g_qr = ts * (-g[0, 1] * qk + g[0, 2] * qj + g[1, 0] * qk - g[1, 2] * qi - g[2, 0] * qj + g[2, 1] * qi)
    g_qi = ts * (g[0, 1] * qj + g[0, 2] * qk + g[1, 0] * qj - 2 * g[1, 1] * qi - g[1, 2] * qr + g[2, 0] * qk + g[2, 1] * qr - 2 * g[2, 2] * qi)
    g_qj = ts * (-2 * g[0, 0] * qj + g[0, 1] * qi + g[0, 2] * qr + g[1, 0] * qi + g[1, 2] * qk - g[2, 0] * qr + g[2, 1] * qk - 2 * g[2, 2] * qj)
    g_qk = ts * (-2 * g[0, 0] * qk - g[0, 1] * qr + g[0, 2] * qi + g[1, 0] * qr - 2 * g[1, 1] * qk + g[1, 2] * qj + g[2, 0] * qi + g[2, 1] * qj)
    c = -g_ts * ts * ts
    g_qr, g_qi, g_qj, g_qk = g_qr + c * qr, g_qi + c * qi, g_qj + c * qj, g_qk + c * qk
    g_k = g_qi * aa[0] + g_qj * aa[1] + g_qk * aa[2]
    g_aa = k * torch.stack([g_qi, g_qj, g_qk])
    dk = (-ang / 24) if small else (0.5 * torch.cos(h) * ang - torch.sin(h)) / (ang * ang)
    g_ang = -0.5 * torch.sin(h) * g_qr + dk * g_k
    if float(ang) > 0:
        g_aa = g_aa + g_ang * aa / ang
    return g_aa.to(dt)


def object_server_bwd(rot, trans, scene_scale, obj_scale, denorm_mat, pts_cano, g_verts=None, g_tfs=None):
    """Reverse mode of oracle.object_server.  Returns g_rot [B,3], g_trans [B,3], g_scene_scale [B], g_obj_scale []."""
    B = rot.shape[0]
    dt = rot.dtype
    D = denorm_mat.to(dt)
    xh = torch.cat([pts_cano.to(dt), torch.ones(pts_cano.shape[0], 1, dtype=dt)], 1)   # [Nv,4]
    g_rot, g_trans, g_ss = torch.zeros(B, 3, dtype=dt), torch.zeros(B, 3, dtype=dt), torch.zeros(B, dtype=dt)
    g_os = torch.zeros((), dtype=dt)
    for b in range(B):
        R = O.axis_angle_to_matrix(rot[b][None])[0].to(dt)
        s = scene_scale[b]
        M = torch.eye(4, dtype=dt)
        M[:3, :3] = s * obj_scale * R
        M[:3, 3] = s * trans[b]
        T = M @ D
        gT = torch.zeros(4, 4, dtype=dt) if g_tfs is None else g_tfs[b].clone()
        if g_verts is not None:
            o = xh @ T.T                                        # [Nv,4]
            g_o = torch.zeros_like(o)
            g_o[:, :3] = g_verts[b] / o[:, 3:4]
            g_o[:, 3] = -(g_verts[b] * o[:, :3]).sum(1) / (o[:, 3] ** 2)
            gT = gT + g_o.T @ xh
        gM = gT @ D.T
        gR = s * obj_scale * gM[:3, :3]
        g_ss[b] = obj_scale * (R * gM[:3, :3]).sum() + (trans[b] * gM[:3, 3]).sum()
        g_os = g_os + s * (R * gM[:3, :3]).sum()
        g_trans[b] = s * gM[:3, 3]
        g_rot[b] = axis_angle_to_matrix_bwd(rot[b], gR)
    return g_rot, g_trans, g_ss, g_os
