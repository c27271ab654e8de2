"""Per-frame SMPL parameter table: host-side mirror of reference `models/body_model_params.py:5-66`.

One `nn.Embedding` per SMPL parameter (betas: a single shared row; global_orient / transl / body_pose:
one row per training frame), zero-initialised and frozen until `init_parameters` /
`set_requires_grad`; `forward(frame_ids)` looks the rows up.  Same attribute names and state-dict
keys (`<name>.weight`) as the reference, so its checkpoints load unchanged.  When the rows require
a gradient (`optim_body_params`, train.py:141-145) the rendering path builds the per-frame tables with
the differentiable builder and the gradients reach these embeddings (SURVEY §8e caveat)."""
import torch
import torch.nn as nn


class BodyModelParams(nn.Module):
    def __init__(self, num_frames, model_type="smpl"):
        super().__init__()
        if model_type != "smpl":
            raise NotImplementedError("only model_type='smpl' is built (every shipped config of the reference)")
        self.num_frames, self.model_type = num_frames, model_type
        self.params_dim = {"betas": 10, "global_orient": 3, "transl": 3, "body_pose": 69}
        self.param_names = self.params_dim.keys()
        for name, dim in self.params_dim.items():
            emb = nn.Embedding(1 if name == "betas" else num_frames, dim)
            emb.weight.data.fill_(0)
            emb.weight.requires_grad = False
            setattr(self, name, emb)

    def init_parameters(self, param_name, data, requires_grad=False):
        if param_name == "betas":
            data = torch.mean(data, dim=0, keepdim=True)
        w = getattr(self, param_name).weight
        w.data = data[..., :self.params_dim[param_name]].to(w.device)
        w.requires_grad = requires_grad

    def set_requires_grad(self, param_name, requires_grad=True):
        getattr(self, param_name).weight.requires_grad = requires_grad

    def forward(self, frame_ids):
        # The rows an nn.Embedding lookup returns (`emb(ids)`, as the reference writes it), read with one elementwise
        # gather per table: the embedding kernel for a handful of indices is serial (7-14 us per table and step);
        # betas is a single shared row: an expand, no launch.  Gradients: scatter-add into the rows / a sum over the batch.
        out = {}
        flat = frame_ids.reshape(-1)
        for name in self.param_names:
            w = getattr(self, name).weight
            if name == "betas":
                rows = w[0].expand(flat.shape[0], -1)
            else:
                rows = torch.gather(w, 0, flat[:, None].expand(-1, w.shape[1]))
            out[name] = rows.reshape(*frame_ids.shape, w.shape[1])
        return out
