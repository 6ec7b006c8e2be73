"""CPU port of the reference's forward path in stock torch ops — TEST/BENCH INFRASTRUCTURE ONLY.

This is "the reference's CPU torch_scatter / SIR path" of BASELINE.json's north_star: the same frame as
fullysparsefusion_b200.fsf.FSF, executed with the ATen CPU operators the reference's Python calls
(torch.unique, index_add_/scatter_reduce_ in place of the absent torch_scatter 2.0.2, F.grid_sample on
the float-cast id planes, nn.Linear/LayerNorm/BatchNorm1d, scipy connected_components on a dense
distance matrix) and, for the un-vendored spconv, the classic gather → matmul → scatter-add per kernel
offset.  Multi-threaded through torch.set_num_threads.  bench.py times it as `cpu_baseline` and as
`--impl reference` (kind "port": the reference itself cannot be installed here — no mmcv / mmdet3d fork /
spconv / torch_scatter).  It reads its weights from an FSF module instance (parameter containers only).
Stage names match fullysparsefusion_b200/fsf.py.  Only tests/ and bench.py may import this module.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from scipy.sparse.csgraph import connected_components
from torch import nn


def seq(mlp, x):
    """Run a build_mlp stack through the stock nn.Sequential path."""
    return nn.Sequential.forward(mlp, x)


def scatter_v2(feat, coors, mode, unq=None):
    """sst_ops.scatter_v2 (:150-177) with torch_scatter replaced by ATen CPU ops."""
    if unq is None:
        new_coors, inv = torch.unique(coors, return_inverse=True, dim=0)
    else:
        new_coors, inv = unq
    m = new_coors.size(0)
    if mode == "max":
        out = torch.full((m, feat.size(1)), float("-inf"), dtype=feat.dtype)
        out.scatter_reduce_(0, inv[:, None].expand_as(feat), feat, reduce="amax", include_self=True)
    else:
        out = torch.zeros((m, feat.size(1)), dtype=feat.dtype).index_add_(0, inv, feat)
        if mode == "avg":
            out = out / torch.bincount(inv, minlength=m).clamp(min=1).to(feat.dtype)[:, None]
    return out, new_coors, inv


def floor_coors(points, voxel, rng, kernel_rule=False):
    """kernel_rule=False: the in-tree formula torch.div(p - min, vs, rounding_mode='floor') (single_stage_fsd.py:270, 444, 591-593,
    948).  kernel_rule=True: floor((p - min) / vs) in fp32, the rule of mmdet3d's dynamic Voxelization kernel behind
    VoteSegmentor.voxelize (single_stage_fsd.py:206-226) — the two differ for a few points per 100 k that sit on a voxel face."""
    lo = torch.tensor(rng[:3], dtype=torch.float32)
    vs = torch.tensor(voxel, dtype=torch.float32)
    if kernel_rule:
        return torch.floor((points[:, :3] - lo[None]) / vs[None]).long()
    return torch.div(points[:, :3] - lo[None], vs[None], rounding_mode="floor").long()


# ---- projection + sampling, literal FSF.py:169-226 ---------------------------------------------------
def prj_points_2d(points, lidar2img, img_h, img_w):
    pts_4d = torch.cat([points[:, :3], points.new_ones((points.size(0), 1))], dim=-1)
    pts_2d = pts_4d @ lidar2img.permute(0, 2, 1)
    depth_valid = pts_2d[..., 2] > 1e-3
    pts_2d[..., 2] = torch.clamp(pts_2d[..., 2], min=1e-5, max=1e5)
    pts_2d[..., 0] /= pts_2d[..., 2]
    pts_2d[..., 1] /= pts_2d[..., 2]
    pts_2d[..., 0] /= img_w
    pts_2d[..., 1] /= img_h
    pts_2d = (pts_2d[..., :2] - 0.5) * 2
    valid = depth_valid & (pts_2d[..., 0] > -1) & (pts_2d[..., 0] < 1) & (pts_2d[..., 1] > -1) & (pts_2d[..., 1] < 1)
    pts_2d[~valid] = -2.0
    return pts_2d


def points_in_mask(points, mask_data, lidar2img):
    cams, classes, H, W = mask_data.shape
    pts_2d = prj_points_2d(points, lidar2img, H, W)
    mask_f = mask_data.float()                                                     # FSF.py:209
    out = []
    for cam in range(cams):
        s = F.grid_sample(mask_f[cam][None], pts_2d[cam][None, None], mode="nearest", align_corners=False)
        out.append(s[0, :, 0, :].permute(1, 0))
    return torch.stack(out, 1).long()


# ---- sparse convolution: gather → matmul → scatter-add per offset (spconv's CPU algorithm) -----------
def _keys(coors, shape):
    k = coors[:, 0].long()
    for a in range(3):
        k = k * shape[1 + a] + coors[:, 1 + a].long()
    return k


def rulebook(out_coors, in_coors, in_shape, stride, pad, transposed=False):
    in_keys = _keys(in_coors, in_shape)            # ascending (rows are in lexicographic order)
    pairs = []
    for kz in range(3):
        for ky in range(3):
            for kx in range(3):
                k = torch.tensor([kz, ky, kx])
                if not transposed:
                    c = out_coors[:, 1:].long() * torch.tensor(stride) - torch.tensor(pad) + k
                    ok = torch.ones(len(c), dtype=torch.bool)
                else:
                    t = out_coors[:, 1:].long() + torch.tensor(pad) - k
                    ok = ((t >= 0) & (t % torch.tensor(stride) == 0)).all(1)
                    c = torch.div(t, torch.tensor(stride), rounding_mode="floor")
                ok &= ((c >= 0) & (c < torch.tensor(in_shape[1:]))).all(1)
                q = _keys(torch.cat([out_coors[:, :1].long(), c.clamp(min=0)], 1), in_shape)
                pos = torch.searchsorted(in_keys, q).clamp(max=max(len(in_keys) - 1, 0))
                hit = ok & (in_keys[pos] == q)
                o = torch.nonzero(hit)[:, 0]
                pairs.append((pos[o], o))
    return pairs


def sparse_conv(x, pairs, n_out, module, residual=None):
    out = torch.zeros((n_out, module.weight.size(1)), dtype=x.dtype)
    for k, (i, o) in enumerate(pairs):
        if len(i):
            out.index_add_(0, o, x[i] @ module.weight[k].t())
    bn = module.bn
    out = F.batch_norm(out, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)
    if residual is not None:
        out = out + residual
    return F.relu(out) if module.act else out


# ---- query refinement + final boxes (FSF.py:960-1095, frustum_cluster_head.py:595-698) on the host ----------------
def decode_boxes(reg, base):
    """BasePointBBoxCoder.decode + the batch column of decode_stage_bboxes."""
    dims = reg[:, 3:6].exp() - 1e-6
    yaw = torch.atan2(reg[:, 6:7], reg[:, 7:8])
    return torch.cat([reg.new_zeros((reg.size(0), 1)), reg[:, :3] + base[:, :3], dims, yaw, reg[:, 8:10]], 1)


def dynamic_point_pool(rois7, pts, extra, max_inbox, capacity, chunk=64):
    """dynamic_point_pool_ext.forward in canonical (roi, point) order (include/fsf_b200.h): brute force over roi chunks."""
    out_p, out_r, out_f = [], [], []
    total = 0
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    for r0 in range(0, rois7.size(0), chunk):
        r = rois7[r0:r0 + chunk]
        rot = r[:, 6:7] + torch.tensor(np.pi / 2, dtype=torch.float32)   # mmdet3d 0.x: rot_angle = rz + pi / 2
        cosa, sina = torch.cos(rot), torch.sin(rot)
        sx, sy, lz = x[None] - r[:, 0:1], y[None] - r[:, 1:2], z[None] - r[:, 2:3]
        lx = sx * cosa + sy * (-sina)
        ly = sx * sina + sy * cosa
        hl, hw, hh = r[:, 4:5] * 0.5, r[:, 3:4] * 0.5, r[:, 5:6] * 0.5
        el, ew, eh = (r[:, 4:5] + extra[0]) * 0.5, (r[:, 3:4] + extra[1]) * 0.5, (r[:, 5:6] + extra[2]) * 0.5
        hit = (lz.abs() <= eh) & (lx > -el) & (lx < el) & (ly > -ew) & (ly < ew)
        for j in range(r.size(0)):
            idx = torch.nonzero(hit[j])[:, 0][:max_inbox]
            idx = idx[: max(0, capacity - total)]
            if idx.numel() == 0:
                continue
            total += idx.numel()
            a, b, c = lx[j, idx], ly[j, idx], lz[j, idx]
            inner = (a.abs() < hl[j]) & (b.abs() < hw[j]) & (c.abs() <= hh[j])
            out_f.append(torch.stack([x[idx], y[idx], z[idx], a, b, c, a + hl[j], b + hw[j], c + hh[j], hl[j] - a, hw[j] - b, hh[j] - c,
                                      (~inner).float()], 1))
            out_p.append(idx)
            out_r.append(torch.full_like(idx, r0 + j))
    if not out_p:
        return (torch.full((1,), -1, dtype=torch.long), torch.full((1,), -1, dtype=torch.long), torch.zeros((1, 13)))
    return torch.cat(out_p), torch.cat(out_r), torch.cat(out_f)


def _corners(b):
    """[n,4,2] counter-clockwise corners of BEV rectangles (x, y, dx, dy, yaw) in float64."""
    c, s = np.cos(b[:, 4]), -np.sin(b[:, 4])      # clockwise by yaw (mmdet3d 0.x iou3d)
    sx = np.stack([b[:, 2], -b[:, 2], -b[:, 2], b[:, 2]], 1) * 0.5
    sy = np.stack([b[:, 3], b[:, 3], -b[:, 3], -b[:, 3]], 1) * 0.5
    return np.stack([b[:, 0:1] + sx * c[:, None] - sy * s[:, None], b[:, 1:2] + sx * s[:, None] + sy * c[:, None]], 2)


def rotated_iou_pairs(a, b):
    """IoU of rectangle pairs a[i] / b[i] ([P,5] = x, y, dx, dy, yaw): Sutherland-Hodgman clipping vectorised over the
    pairs (polygons of at most 8 vertices in fixed-size buffers with a per-pair vertex count)."""
    P = a.shape[0]
    poly = np.zeros((P, 8, 2))
    poly[:, :4] = _corners(a)
    cnt = np.full(P, 4)
    clip = _corners(b)
    ar = np.arange(P)
    for e in range(4):
        p0, edge = clip[:, e], clip[:, (e + 1) % 4] - clip[:, e]
        new = np.zeros((P, 8, 2))
        ncnt = np.zeros(P, dtype=np.int64)
        for j in range(8):
            live = j < cnt
            q0 = poly[:, j]
            q1 = poly[ar, np.where(j + 1 < cnt, j + 1, 0)]
            d0 = edge[:, 0] * (q0[:, 1] - p0[:, 1]) - edge[:, 1] * (q0[:, 0] - p0[:, 0])
            d1 = edge[:, 0] * (q1[:, 1] - p0[:, 1]) - edge[:, 1] * (q1[:, 0] - p0[:, 0])
            keep = live & (d0 >= 0)
            sel = np.flatnonzero(keep & (ncnt < 8))
            new[sel, ncnt[sel]] = q0[sel]
            ncnt[sel] += 1
            cross = live & ((d0 >= 0) != (d1 >= 0))
            sel = np.flatnonzero(cross & (ncnt < 8))
            t = d0[sel] / (d0[sel] - d1[sel])
            new[sel, ncnt[sel]] = q0[sel] + (q1[sel] - q0[sel]) * t[:, None]
            ncnt[sel] += 1
        poly, cnt = new, ncnt
    nxt = np.where(np.arange(8)[None] + 1 < cnt[:, None], np.arange(8)[None] + 1, 0)
    x, y = poly[:, :, 0], poly[:, :, 1]
    xn, yn = np.take_along_axis(x, nxt, 1), np.take_along_axis(y, nxt, 1)
    valid = np.arange(8)[None] < cnt[:, None]
    inter = 0.5 * np.abs(((x * yn - xn * y) * valid).sum(1))
    return inter / np.maximum(a[:, 2] * a[:, 3] + b[:, 2] * b[:, 3] - inter, 1e-8)


def multiclass_nms(boxes, logits, score_thr, nms_thr, max_num):
    """box3d_multiclass_nms with rotated BEV IoU on sigmoid(logits): candidate pairs pre-filtered by centre distance, IoUs
    vectorised, greedy pass per class in (score desc, box index asc) order."""
    scores = torch.sigmoid(logits).numpy()
    bx = boxes.numpy().astype(np.float64)
    bev = bx[:, [0, 1, 3, 4, 6]]
    rad = 0.5 * np.hypot(bev[:, 2], bev[:, 3])
    rows, sc, lb = [], [], []
    for c in range(scores.shape[1]):
        idx = np.flatnonzero(scores[:, c] > np.float32(score_thr))
        if idx.size == 0:
            continue
        idx = idx[np.lexsort((idx, -scores[idx, c].astype(np.float64)))]
        b = bev[idx]
        n = idx.size
        sup = [[] for _ in range(n)]            # sup[i]: later boxes that i suppresses
        for r0 in range(0, n, 512):
            d = np.hypot(b[r0:r0 + 512, None, 0] - b[None, :, 0], b[r0:r0 + 512, None, 1] - b[None, :, 1])
            near = d < (rad[idx][r0:r0 + 512, None] + rad[idx][None, :])
            ii, jj = np.nonzero(near)
            ii += r0
            m = jj > ii
            ii, jj = ii[m], jj[m]
            if ii.size:
                hitp = rotated_iou_pairs(b[ii], b[jj]) > nms_thr
                for i, j in zip(ii[hitp], jj[hitp]):
                    sup[i].append(j)
        removed = np.zeros(n, bool)
        for i in range(n):
            if not removed[i]:
                rows.append(idx[i])
                sc.append(scores[idx[i], c])
                lb.append(c)
                if sup[i]:
                    removed[sup[i]] = True
    rows, sc, lb = np.asarray(rows, np.int64), np.asarray(sc, np.float32), np.asarray(lb, np.int64)
    if rows.size > max_num:
        order = np.lexsort((np.arange(rows.size), -sc.astype(np.float64)))[:max_num]
        rows, sc, lb = rows[order], sc[order], lb[order]
    return boxes[torch.from_numpy(rows)], torch.from_numpy(sc), torch.from_numpy(lb), torch.from_numpy(rows)


class CpuFSF:
    def __init__(self, model):
        self.m = model
        self.cfg = model.cfg

    def stages(self, points, mask_data, mask_anno, lidar2img) -> Tuple[List[Tuple[str, Callable[[], None]]], Dict]:
        m, cfg, st = self.m, self.cfg, {}
        rng = cfg["point_cloud_range"]
        P = int(cfg.get("point_dim", 5))          # nuScenes: x,y,z,intensity,dt; AV2: x,y,z,intensity; then the un-augmented xyz
        is_argo = bool(cfg.get("is_argo", False))

        def sir_layer(layer, feats, inv, unq, f_cluster, return_pts=True):
            x = torch.cat([feats[:, :3] / torch.tensor(layer.xyz_normalizer), feats[:, 3:]], 1)
            x = x * seq(layer.rel_mlp, f_cluster / layer.rel_dist_scaler)
            ori, cl = x, []
            for i, vfe in enumerate(layer.vfe_layers):
                pf = F.gelu(vfe.norm(vfe.linear(x))) if vfe.act == "gelu" else F.relu(vfe.norm(vfe.linear(x)))
                c, _, _ = scatter_v2(pf, None, "max", unq=(unq, inv))
                cl.append(c)
                if i != len(layer.vfe_layers) - 1:
                    x = torch.cat([pf, c[inv]], 1)
            if pf.shape == ori.shape:
                pf = pf + ori
            return pf, torch.cat(cl, 1)

        def sir(net, pts, feats, coors, f_cluster):
            unq, inv = torch.unique(coors, return_inverse=True, dim=0)
            out, cl = feats, []
            for block in net.block_list:
                out, c = sir_layer(block, torch.cat([pts, out], 1), inv, unq, f_cluster)
                cl.append(c)
            return out, torch.cat(cl, 1), unq

        def head(h, x):
            x = seq(h.shared_mlp, x)
            r = {k: seq(getattr(h.task_heads[0], k), x) for k in h.task_heads[0].attrs}
            return r["score"], torch.cat([r["center"], r["dim"], r["rot"]] + ([r["vel"]] if "vel" in r else []), 1)

        def segment():
            pts5 = points[:, :P]
            c = floor_coors(points, cfg["seg_voxel_size"], rng, kernel_rule=True)[:, [2, 1, 0]]
            coors = F.pad(c, (1, 0), value=0)
            unq, inv = torch.unique(coors, return_inverse=True, dim=0)
            vfe = m.voxel_encoder
            mean, _, _ = scatter_v2(pts5, None, "avg", unq=(unq, inv))
            f_cluster = pts5[:, :3] - mean[inv][:, :3]
            vs = torch.tensor(cfg["seg_voxel_size"])
            off = torch.tensor([cfg["seg_voxel_size"][a] / 2 + rng[a] for a in range(3)])
            f_center = pts5[:, :3] - (coors[:, [3, 2, 1]].float() * vs + off)
            x = torch.cat([pts5, f_cluster, f_center], 1)
            for i, layer in enumerate(vfe.vfe_layers):
                pf = nn.Sequential.forward(layer, x)
                vf, _, _ = scatter_v2(pf, None, "max", unq=(unq, inv))
                if i != len(vfe.vfe_layers) - 1:
                    x = torch.cat([pf, vf[inv]], 1)
            # SimpleSparseUNet
            net = m.backbone_unet
            shape = [1] + list(cfg["sparse_shape"])
            levels = [(unq, shape)]
            rb = {"subm1": rulebook(unq, unq, shape, [1, 1, 1], [1, 1, 1])}
            for i in range(1, net.stage_num):
                pad = net._triple(tuple(net.encoder_paddings[i])[0])
                pc, ps = levels[-1]
                oshape = [1] + [(ps[1 + a] + 2 * pad[a] - 3) // 2 + 1 for a in range(3)]
                cand = []
                for kz in range(3):
                    for ky in range(3):
                        for kx in range(3):
                            t = pc[:, 1:] + torch.tensor(pad) - torch.tensor([kz, ky, kx])
                            ok = ((t >= 0) & (t % 2 == 0)).all(1)
                            o = torch.div(t, 2, rounding_mode="floor")
                            ok &= (o < torch.tensor(oshape[1:])).all(1)
                            cand.append(torch.cat([pc[ok, :1], o[ok]], 1))
                oc = torch.unique(torch.cat(cand), dim=0)
                rb[f"spconv{i + 1}"] = rulebook(oc, pc, ps, [2, 2, 2], pad)
                rb[f"spconv{i + 1}_inv"] = rulebook(pc, oc, oshape, [2, 2, 2], pad, transposed=True)
                rb[f"subm{i + 1}"] = rulebook(oc, oc, oshape, [1, 1, 1], [1, 1, 1])
                levels.append((oc, oshape))
            x = sparse_conv(vf, rb["subm1"], len(unq), net.conv_input)
            enc = []
            for i, stage in enumerate(net.encoder_layers):
                for layer in stage:
                    x = sparse_conv(x, rb[layer.indice_key], len(levels[i][0]), layer)
                enc.append(x)
            x = enc[-1]
            for lvl in range(net.stage_num, 0, -1):
                lat_in, n_l = enc[lvl - 1], len(levels[lvl - 1][0])
                lat = getattr(net, f"lateral_layer{lvl}")
                h = sparse_conv(lat_in, rb[f"subm{lvl}"], n_l, lat.conv1)
                h = sparse_conv(h, rb[f"subm{lvl}"], n_l, lat.conv2, residual=lat_in)
                cat = torch.cat([x, h], 1)
                merged = sparse_conv(cat, rb[f"subm{lvl}"], n_l, getattr(net, f"merge_layer{lvl}"))
                x = merged + cat.view(n_l, merged.size(1), -1).sum(2)
                up = getattr(net, f"upsample_layer{lvl}")
                x = sparse_conv(x, rb[f"spconv{lvl}_inv"] if lvl != 1 else rb["subm1"], len(levels[max(lvl - 2, 0)][0]), up)
            # Voxel2PointScatterNeck
            pts_feats = x[inv]
            centre = (coors[:, [3, 2, 1]].float() + 0.5) * vs + torch.tensor(rng[:3])
            st.update(voxel_feats=x, pts_lidar_feats=torch.cat([pts_feats, pts5[:, :3] - centre], 1))

        def enhance():
            ids = points_in_mask(points[:, P:P + 3], mask_data, lidar2img)           # frustum_gather (FSF.py:228-258)
            cam = ids.sum(-1).max(-1)[1]
            sel = F.one_hot(cam, ids.size(1)).bool().unsqueeze(-1)
            ids_sel = ids.masked_select(sel).reshape(-1, ids.size(2))
            preds = torch.zeros((len(ids_sel), ids.size(2), mask_anno.size(1)))
            valid = ids_sel >= 1
            preds[valid] = mask_anno[ids_sel[valid] - 1]                             # get_all_cls_preds_2d
            if is_argo:   # encode_2d_feats with encode_single_cls (FSF.py:449-474, 537-552): bbox / w,h ‖ score ‖ one-hot category
                pr = preds.reshape(-1, mask_anno.size(1)).clone()
                pr[~valid.reshape(-1), 5] = m.num_classes                            # invalid rows: category = "none"
                box = pr[:, :4].clone()
                box[:, 0::2] /= mask_data.shape[-1]
                box[:, 1::2] /= mask_data.shape[-2]
                pt2d = torch.cat([box, pr[:, 4:5], F.one_hot(pr[:, 5].long(), m.num_classes + 1).float()], 1)
            else:
                pt2d = preds[..., 4]
            img_feat = seq(m.segmentor_updated_mlp, pt2d)
            st["img_scores"] = pt2d
            feats = st["pts_lidar_feats"] + img_feat
            h = seq(m.segmentation_head.pre_seg_conv, feats)
            logits, votes = m.segmentation_head.conv_seg(h), m.segmentation_head.voting(h)
            st.update(ids=ids, seg_feats=feats, seg_logits=logits, seg_vote_preds=votes, offsets=votes * votes.abs())

        def frustum():
            pts5, ids = points[:, :P], st["ids"]
            fgw = 1 - st["seg_logits"].softmax(1)[:, -1]
            fg = ids.sum((-2, -1)) > 0                                              # extract_fg_pts
            feat, p, o, w = st["seg_feats"][fg], pts5[fg], ids[fg].reshape(int(fg.sum()), -1), fgw[fg]
            ov = (o > 0).sum(-1)
            raw = o.max(-1)[0]
            feat_c, p_c, w_c = feat.clone(), p.clone(), w.clone()
            for k in range(2, int(ov.max()) + 1 if len(ov) else 0):                  # double_overlap_pts
                mk = ov == k
                if mk.sum() == 0:
                    continue
                feat = torch.cat([feat, feat_c[mk].repeat(k - 1, 1)])
                p = torch.cat([p, p_c[mk].repeat(k - 1, 1)])
                w = torch.cat([w, w_c[mk].repeat(k - 1)])
                sv = o[mk].topk(k, dim=-1)[0]
                for pad in range(1, k):
                    raw = torch.cat([raw, sv[:, pad]])
            sir_coors = torch.stack([torch.zeros_like(raw), torch.zeros_like(raw), raw], 1)
            wc = w.clamp(min=1e-5)[:, None]
            mean, _, inv = scatter_v2(torch.cat([p[:, :3] * wc, wc], 1), sir_coors, "avg")
            center = mean[:, :3] / mean[:, 3:4]
            _, cl, oc = sir(m.frustum_sir, p, feat, sir_coors, p[:, :3] - center[inv])
            pr = torch.zeros((len(oc), mask_anno.size(1)))
            ok = oc[:, 2] >= 1
            pr[ok] = mask_anno[oc[ok, 2] - 1]
            pr[~ok, 5] = m.num_classes
            box = pr[:, :4].clone()
            box[:, 0::2] /= mask_data.shape[-1]
            box[:, 1::2] /= mask_data.shape[-2]
            enc = torch.cat([box, pr[:, 4:5], F.one_hot(pr[:, 5].long(), m.num_classes + 1).float()], 1)
            obj = torch.cat([cl, seq(m.encode_2d_mlp, enc)], 1)
            st.update(frustum_obj_feats=obj, frustum_out=head(m.frustum_obj_head, obj), frustum_centers=center)

        def fsd():
            pts5 = points[:, :P]
            c = F.pad(floor_coors(pts5, cfg["pre_voxelization_size"], rng)[:, [2, 1, 0]], (1, 0), value=0)
            unq = torch.unique(c, return_inverse=True, dim=0)
            v = {k: scatter_v2(t, None, "avg", unq=unq)[0] for k, t in dict(p=pts5, l=st["seg_logits"], v=st["seg_vote_preds"],
                                                                            f=st["seg_feats"], o=st["offsets"]).items()}
            scores = v["l"].softmax(1)
            off = v["o"].reshape(-1, m.num_classes + 1, 3)
            rows, inds, ctrs = [], [], []
            for g, idx in enumerate(m.groups):
                fgm = scores[:, idx].sum(1) > cfg["score_thresh"][g]
                if not fgm.any():
                    fgm[0] = True
                lg = v["l"][:, idx][fgm]
                wt = ((lg - lg.max(1)[0][:, None]).abs() < 1e-6).float()
                wt = wt / wt.sum(1)[:, None]
                ctr = v["p"][fgm, :3] + (off[:, idx, :][fgm] * wt[:, :, None]).sum(1)
                cc = F.pad(floor_coors(ctr, cfg["cluster_voxel_size"][g], rng), (1, 0), value=0)   # forward_single_class
                _, inv, cnt = torch.unique(cc, return_inverse=True, return_counts=True, dim=0)
                valid = cnt[inv] >= cfg["min_points"]
                if not valid.any():
                    valid = ~valid
                sc, _, inv2 = scatter_v2(ctr[valid], cc[valid], "avg")
                d = ((sc[:, None, :2] - sc[None, :, :2]) ** 2).sum(2) ** 0.5          # find_connected_componets_single_batch
                lab = torch.from_numpy(connected_components((d < cfg["connected_dist"][g]).numpy(), directed=False)[1]).long()
                rows.append(torch.nonzero(fgm)[:, 0][valid])
                inds.append(torch.stack([torch.full_like(lab[inv2], g), torch.zeros_like(lab[inv2]), lab[inv2]], 1))
                ctrs.append(ctr[valid])
            rows, inds, ctrs = torch.cat(rows), torch.cat(inds), torch.cat(ctrs)
            feats = torch.cat([v["l"][rows], v["v"][rows], v["f"][rows]], 1)
            cxyz, _, inv = scatter_v2(ctrs, inds, "avg")
            _, cl, _ = sir(m.backbone, v["p"][rows], feats, inds, v["p"][rows, :3] - cxyz[inv])
            st.update(fsd_obj_feats=cl, fsd_out=head(m.bbox_head, cl), fsd_rows=rows, fsd_centers=cxyz)

        def combine():
            st["obj_feats"] = torch.cat([seq(m.combine_frustum_feat_mlp, st["frustum_obj_feats"]),
                                         seq(m.combine_fsd_feat_mlp, st["fsd_obj_feats"])], 0)

        def refine():
            # each_stage_refine / query_feat_refine (FSF.py:1009-1083), one extra stage
            pts5 = points[:, :P]
            centers = torch.cat([st["frustum_centers"], st["fsd_centers"]], 0)
            reg = torch.cat([st["frustum_out"][1], st["fsd_out"][1]], 0)
            res = st["obj_feats"]
            for i in range(m.num_extra_stages):
                rois = decode_boxes(reg, centers)
                centers = rois[:, 1:4]
                inds, roi_inds, info = dynamic_point_pool(rois[:, 1:8], pts5[:, :3], m.roi_extractor.extra_wlh, m.roi_extractor.max_inbox_point,
                                                          m.roi_extractor.max_all_pts)
                ex_pts, ex_feats = pts5[inds], st["seg_feats"][inds]
                feats = torch.cat([ex_feats, seq(m.refine_img_mlp[i], st["img_scores"][inds])], 1)
                head_i = m.refine_sir_layers[i]
                rel_xyz = ex_pts[:, :3] - centers[roi_inds]
                f_cluster = torch.cat([info[:, 3:6], info[:, 6:12], info[:, 12:13], rel_xyz], 1)
                unq, inv = torch.unique(roi_inds[:, None], return_inverse=True, dim=0)
                out, cl = feats, []
                for block in head_i.block_list:
                    out, c = sir_layer(block, torch.cat([ex_pts, out, f_cluster / 10], 1), inv, unq, f_cluster)
                    cl.append(c)
                cl = torch.cat(cl, 1)
                lidar = torch.zeros((rois.size(0), cl.size(1)))
                ok = unq[:, 0] >= 0
                lidar[unq[ok, 0]] = cl[ok]
                query = seq(m.out_proj[i], seq(m.lidar_img_mlp[i], lidar) + res + seq(m.position_encoder[i], centers))
                cls, reg = head(m.frustum_refined_head[i], query)
                res = query
                st.update({f"refine{i}_rois": rois, f"refine{i}_pts_inds": inds, f"refine{i}_roi_inds": roi_inds, f"refine{i}_lidar_feat": lidar,
                           f"refine{i}_query": query, f"refine{i}_cls": cls, f"refine{i}_reg": reg, f"refine{i}_centers": centers})

        def boxes():
            # FrustumClusterHead._get_bboxes_single (frustum_cluster_head.py:595-698): sigmoid, decode, rotated multi-class NMS
            i = m.num_extra_stages - 1
            rois = decode_boxes(st[f"refine{i}_reg"], st[f"refine{i}_centers"])
            b, s_, l_, r_ = multiclass_nms(rois[:, 1:], st[f"refine{i}_cls"], 0.01, 0.35, 500)
            st.update(det_boxes=b, det_scores=s_, det_labels=l_, det_rows=r_)

        self.extra_stages = [("refine", refine), ("boxes", boxes)]
        return [("segment", segment), ("enhance", enhance), ("frustum", frustum), ("fsd", fsd), ("combine", combine)], st
