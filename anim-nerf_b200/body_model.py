"""Per-frame table builder: SMPL linear blend skinning in plain torch (host side).

SURVEY §8 row A16: the per-frame inputs of the hot path (posed vertices, per-vertex
4x4 transforms, shape/pose offsets).  This torch formulation is the differentiable
builder, used when SMPL parameters are being optimised (autograd reaches them); every
other step / frame goes through the fused kernels (`an_body_tables_fwd`,
csrc/body_tables.cu), which read the constant buffers registered here.  The arithmetic follows the reference's *modified*
smplx (`smplx/lbs.py:152-251` `lbs`, `:298-330` `batch_rodrigues`, `:348-420`
`batch_rigid_transform`; `smplx/body_models.py:289-387` `SMPL.forward`, which adds
`transl` into `A` and `T` and returns the shape/pose offsets) so that the tables
agree with the reference's to fp32 round-off.  Written from that description, not
copied: one einsum-based implementation, no SMPL-H/X/MANO/FLAME variants.
"""
import pickle

import numpy as np
import torch
import torch.nn as nn


def rodrigues(rvec):
    """axis-angle (N,3) -> rotation matrices (N,3,3); keeps the reference's
    `norm(r + 1e-8)` angle (smplx/lbs.py:316)."""
    angle = torch.linalg.vector_norm(rvec + 1e-8, dim=1, keepdim=True)
    axis = rvec / angle
    c = torch.cos(angle)[:, :, None]
    s = torch.sin(angle)[:, :, None]
    x, y, z = axis[:, 0], axis[:, 1], axis[:, 2]
    o = torch.zeros_like(x)
    K = torch.stack([o, -z, y, z, o, -x, -y, x, o], dim=1).view(-1, 3, 3)
    eye = torch.eye(3, dtype=rvec.dtype, device=rvec.device)[None]
    return eye + s * K + (1 - c) * torch.bmm(K, K)


def rigid_chain(rot, joints, parents, parent_idx=None):
    """rot (B,J,3,3), rest joints (B,J,3) -> posed joints (B,J,3) and the relative
    transforms A (B,J,4,4) that map rest-pose points to posed points."""
    B, J = joints.shape[:2]
    parents = [int(p) for p in parents]               # host list: no device read-back inside the chain
    rel = joints.clone()
    if parent_idx is None:
        parent_idx = torch.as_tensor(parents[1:], dtype=torch.long, device=joints.device)
    rel[:, 1:] = joints[:, 1:] - joints.index_select(1, parent_idx)
    local = torch.zeros(B, J, 4, 4, dtype=rot.dtype, device=rot.device)
    local[:, :, :3, :3] = rot
    local[:, :, :3, 3] = rel
    local[:, :, 3, 3] = 1.0
    chain = [local[:, 0]]
    for j in range(1, J):
        chain.append(torch.matmul(chain[parents[j]], local[:, j]))
    G = torch.stack(chain, dim=1)                      # world transform of each joint
    posed = G[:, :, :3, 3]
    # subtract G @ [rest joint; 0] from the translation column
    jh = torch.cat([joints, torch.zeros_like(joints[..., :1])], -1)[..., None]  # (B,J,4,1)
    corr = torch.matmul(G, jh)                          # (B,J,4,1)
    A = G - torch.cat([torch.zeros_like(G[..., :3]), corr], dim=-1)
    return posed, A


class BodyModel(nn.Module):
    """SMPL-schema body model.  `forward(betas, body_pose, global_orient, transl)`
    returns the dict the reference's `AnimNeRF.set_body_model` reads
    (models/anim_nerf.py:108-126): vertices, joints (24), joints_transform,
    vertices_transform, shape_offsets, pose_offsets."""

    def __init__(self, data, dtype=torch.float32):
        super().__init__()
        if isinstance(data, str):
            with open(data, "rb") as fh:
                data = pickle.load(fh, encoding="latin1")
        t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype)
        self.register_buffer("v_template", t(data["v_template"]))
        self.register_buffer("shapedirs", t(np.asarray(data["shapedirs"])[:, :, :10]))
        nposes = np.asarray(data["posedirs"]).shape[-1]
        self.register_buffer("posedirs", t(np.reshape(data["posedirs"], [-1, nposes]).T).contiguous())
        self.register_buffer("J_regressor", t(data["J_regressor"]))
        parents = torch.as_tensor(np.asarray(data["kintree_table"])[0]).long().clone()
        parents[0] = -1
        self.register_buffer("parents", parents)
        self.parents_host = [int(p) for p in parents]   # the kinematic tree is static: keep it on the host
        self.register_buffer("parent_idx", parents[1:].clone())
        self.register_buffer("lbs_weights", t(data["weights"]))
        # constants of the fused table builder (an_body_tables_fwd): the joint regressor applied to the
        # linear shape model once, and the kinematic tree as int32
        self.register_buffer("J_template", torch.matmul(self.J_regressor, self.v_template).contiguous())
        self.register_buffer("J_shapedirs", torch.einsum("ji,ikl->jkl", self.J_regressor, self.shapedirs).contiguous())
        self.register_buffer("parents_i32", parents.to(torch.int32))

    def refresh_derived(self):
        """Recompute everything derived from v_template / shapedirs / J_regressor / parents (after `load_state_dict` of a
        reference checkpoint replaced those buffers): the fused table builder's constants and the host-side tree."""
        parents = self.parents.long().clone()
        parents[0] = -1
        self.parents.copy_(parents)
        self.parents_host = [int(p) for p in parents]
        self.parent_idx.copy_(parents[1:])
        self.J_template.copy_(torch.matmul(self.J_regressor, self.v_template))
        self.J_shapedirs.copy_(torch.einsum("ji,ikl->jkl", self.J_regressor, self.shapedirs))
        self.parents_i32.copy_(parents.to(torch.int32))

    def forward(self, betas, body_pose, global_orient, transl=None, **_):
        B = max(betas.shape[0], body_pose.shape[0], global_orient.shape[0])
        if betas.shape[0] != B:
            betas = betas.expand(B, -1)
        pose = torch.cat([global_orient, body_pose], dim=1)
        shape_offsets = torch.einsum("bl,mkl->bmk", betas, self.shapedirs)
        v_shaped = self.v_template[None] + shape_offsets
        J = torch.einsum("bik,ji->bjk", v_shaped, self.J_regressor)
        rot = rodrigues(pose.reshape(-1, 3)).view(B, -1, 3, 3)
        eye = torch.eye(3, dtype=rot.dtype, device=rot.device)
        feat = (rot[:, 1:] - eye).reshape(B, -1)
        pose_offsets = torch.matmul(feat, self.posedirs).view(B, -1, 3)
        v_posed = pose_offsets + v_shaped
        joints, A = rigid_chain(rot, J, self.parents_host, self.parent_idx)
        nj = self.J_regressor.shape[0]
        T = torch.matmul(self.lbs_weights[None].expand(B, -1, -1), A.view(B, nj, 16)).view(B, -1, 4, 4)
        vh = torch.cat([v_posed, torch.ones_like(v_posed[..., :1])], dim=2)
        verts = torch.matmul(T, vh[..., None])[:, :, :3, 0]
        if transl is not None:
            joints = joints + transl[:, None]
            verts = verts + transl[:, None]
            shift = torch.zeros_like(A)
            shift[..., :3, 3] = transl[:, None]
            A = A + shift
            shiftT = torch.zeros_like(T)
            shiftT[..., :3, 3] = transl[:, None]
            T = T + shiftT
        return dict(vertices=verts, joints=joints, joints_transform=A, vertices_transform=T,
                    shape_offsets=shape_offsets, pose_offsets=pose_offsets)
