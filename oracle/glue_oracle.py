"""Oracle (test infrastructure only): the driver-side code around the sampling loops, restated on the CPU.

* inter-round glue                 test_amass_full.py:256-258, 268-311
* get_repr_smplx (trajectory part) data_loaders/motion_representation.py:187-282 (+ quaternion.py qbetween/qmul/qinv/qrot,
                                   utils/other_utils.py:264-279 estimate_angular_velocity_np)
* PoseNet condition + masks        test_amass_full.py:313-370
* 'joint_rel_traj' recovery        data_loaders/motion_representation.py:308-329, 355-371
* 2-D reprojection guidance        model/posenet.py:260-317, utils/other_utils.py:150-185

numpy / torch-CPU, same dtypes as the reference at each step (float32 quaternion helpers, float64 scipy-style rotation
matrices).  Pinned by tests/golden/glue.npz, produced by the unmodified reference functions (tools/gen_golden.py).
"""
import numpy as np
import torch

from . import kinematics_oracle as ko

FACE_JOINTS = (2, 1, 17, 16)  # r_hip, l_hip, sdr_r, sdr_l (motion_representation.py:15)


# ------------------------------------------------------------------------------------------------------------
# small quaternion / rotation helpers (numpy in, numpy out; float32 like the reference's *_np wrappers)
# ------------------------------------------------------------------------------------------------------------
def _qbetween(v0, v1):
    v0, v1 = torch.from_numpy(v0).float(), torch.from_numpy(v1).float()
    v = torch.cross(v0, v1, dim=-1)
    w = torch.sqrt((v0 ** 2).sum(-1, keepdim=True) * (v1 ** 2).sum(-1, keepdim=True)) + (v0 * v1).sum(-1, keepdim=True)
    q = torch.cat([w, v], dim=-1)
    return (q / torch.norm(q, dim=-1, keepdim=True)).numpy()


def _qinv(q):
    q = torch.from_numpy(q).float()
    return (q * torch.tensor([1.0, -1.0, -1.0, -1.0])).numpy()


def _qmul(q, r):
    q, r = torch.from_numpy(q).float(), torch.from_numpy(r).float()
    t = torch.bmm(r.view(-1, 4, 1), q.view(-1, 1, 4))
    w = t[:, 0, 0] - t[:, 1, 1] - t[:, 2, 2] - t[:, 3, 3]
    x = t[:, 0, 1] + t[:, 1, 0] - t[:, 2, 3] + t[:, 3, 2]
    y = t[:, 0, 2] + t[:, 1, 3] + t[:, 2, 0] - t[:, 3, 1]
    z = t[:, 0, 3] - t[:, 1, 2] + t[:, 2, 1] + t[:, 3, 0]
    return torch.stack((w, x, y, z), dim=1).numpy()


def _qrot(q, v):
    q, v = torch.from_numpy(q).float(), torch.from_numpy(v).float()
    return ko.qrot(q, v).numpy()


def rotvec_to_matrix(rv):
    """scipy.spatial.transform.Rotation.from_rotvec(rv).as_matrix() (float64): rotvec -> unit quaternion -> matrix."""
    rv = np.asarray(rv, dtype=np.float64)
    ang = np.linalg.norm(rv, axis=-1)
    small = ang <= 1e-3
    a2 = ang * ang
    scale = np.where(small, 0.5 - a2 / 48 + a2 * a2 / 3840, np.sin(ang / 2) / np.where(small, 1.0, ang))
    x, y, z = scale * rv[..., 0], scale * rv[..., 1], scale * rv[..., 2]
    w = np.cos(ang / 2)
    R = np.empty(rv.shape[:-1] + (3, 3))
    R[..., 0, 0] = x * x - y * y - z * z + w * w
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 1, 1] = -x * x + y * y - z * z + w * w
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 2] = -x * x - y * y + z * z + w * w
    return R


# ------------------------------------------------------------------------------------------------------------
# get_repr_smplx, channels [0, 22)
# ------------------------------------------------------------------------------------------------------------
def traj_repr_from_joints(positions, global_orient_aa, transl):
    """positions [T,22,3] float32, global_orient_aa [T,3], transl [T,3] -> [T-1, 22] (REPR_LIST order, un-normalised)."""
    r_hip, l_hip, sdr_r, sdr_l = FACE_JOINTS  # the reference unpacks them as (l_hip, r_hip, sdr_r, sdr_l) = (2, 1, 17, 16)
    across = (positions[:, l_hip] - positions[:, r_hip]) + (positions[:, sdr_r] - positions[:, sdr_l])
    across = across / np.sqrt((across ** 2).sum(axis=-1))[:, None]
    forward = np.cross(np.array([[0, 0, 1]]), across, axis=-1)
    forward = forward / np.sqrt((forward ** 2).sum(axis=-1))[..., None]
    target = np.array([[0, 1, 0]]).repeat(len(forward), axis=0)
    q = _qbetween(forward, target)
    if np.isnan(q).sum() > 0:
        idx = np.where(np.isnan(q))[0][0]
        q[idx] = q[idx - 1]
    q[0] = np.array([[1.0, 0.0, 0.0, 0.0]])
    q_vel = _qmul(q[1:], _qinv(q[:-1]))
    root_l_pos = positions[:, 0]
    root_l_vel = _qrot(q[1:], (positions[1:, 0] - positions[:-1, 0]).copy())
    ang = np.arctan2(q[:, 3:4], q[:, 0:1])
    ang_vel = np.arctan2(q_vel[:, 3:4], q_vel[:, 0:1])
    R = rotvec_to_matrix(global_orient_aa)
    rot6d = R[..., :-1].reshape(-1, 6)
    w = np.matmul(R[1:] - R[:-1], np.transpose(R[:-1], (0, 2, 1)))
    rot_vel = np.stack([(-w[..., 1, 2] + w[..., 2, 1]) / 2.0, (w[..., 0, 2] - w[..., 2, 0]) / 2.0,
                        (-w[..., 0, 1] + w[..., 1, 0]) / 2.0], axis=-1)
    trans_vel = transl[1:] - transl[:-1]
    return np.concatenate([ang[0:-1], ang_vel, root_l_pos[0:-1, [0, 1]], root_l_vel[:, [0, 1]], positions[:-1, 0, 2:3],
                           rot6d[0:-1], rot_vel, transl[0:-1], trans_vel], axis=-1)


def compose_repr(traj_out, clean, repr_abs_only=True):
    """test_amass_full.py:269-277."""
    if not repr_abs_only:
        return torch.cat([traj_out, clean[:, :, traj_out.shape[-1]:]], dim=-1)
    out = clean.clone()
    out[..., 0] = traj_out[..., 0]
    out[..., 2:4] = traj_out[..., 1:3]
    out[..., 6] = traj_out[..., 3]
    out[..., 7:13] = traj_out[..., 4:10]
    out[..., 16:19] = traj_out[..., 10:13]
    return out


def traj_to_full_repr(traj_out, clean, traj_mean, traj_std, pose_mean, pose_std, body_model, repr_abs_only=True):
    """test_amass_full.py:268-311 -> (composite [B,T,294] normalised, traj_rec_full [B,T-1,22] normalised)."""
    composite = compose_repr(traj_out, clean, repr_abs_only)
    full = composite.numpy() * traj_std + traj_mean
    rep = ko.split_repr(torch.from_numpy(full))
    joints = ko.joints_from_smplx(rep, body_model).numpy()
    out = []
    for i in range(joints.shape[0]):
        go = ko.rotmat_to_aa(ko.rot6d_to_rotmat(rep['smplx_rot_6d'][i])).numpy()
        r = traj_repr_from_joints(joints[i], go, rep['smplx_trans'][i].numpy())
        out.append((r - pose_mean[0:22]) / pose_std[0:22])
    return composite, torch.tensor(np.asarray(out)).float()


def pose_to_control_cond(pose_out, T, pose_feat_dim=272):
    """test_amass_full.py:256-258."""
    cc = torch.zeros(pose_out.shape[0], T, pose_feat_dim)
    cc[:, 0:-1] = pose_out[:, :, 0].permute(0, 2, 1)[:, :, -pose_feat_dim:]
    cc[:, -1] = cc[:, -2].clone()
    return cc


def build_pose_cond(src_cl, traj_full, mask_scheme, apply_mask, start=None, end=None, traj_feat_dim=22):
    """test_amass_full.py:333-370 on a channels-last source [B,Tp,294] -> [B,294,1,Tp]."""
    cond = src_cl.clone()
    if traj_full is not None:
        cond[:, :, 0:22] = traj_full
    if apply_mask:
        if mask_scheme in ('lower', 'upper'):
            ids = np.asarray([1, 2, 4, 5, 7, 8, 10, 11] if mask_scheme == 'lower'
                             else [3, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19, 20])
            for k in range(3):
                cond[:, :, traj_feat_dim + ids * 3 + k] = 0.
                cond[:, :, traj_feat_dim + 22 * 3 + ids * 3 + k] = 0.
            for k in range(6):
                cond[:, :, traj_feat_dim + 22 * 3 + 22 * 3 + (ids - 1) * 6 + k] = 0.
            cond[:, :, -4:] = 0.
        elif mask_scheme == 'full':
            cond[:, :, -4:] = 0.
            for i in range(cond.shape[0]):
                cond[i, int(start[i]):int(end[i]), 22:] = 0
    return torch.permute(cond, (0, 2, 1)).unsqueeze(-2).contiguous()


# ------------------------------------------------------------------------------------------------------------
# joint_rel_traj recovery
# ------------------------------------------------------------------------------------------------------------
def joints_from_rel_traj(rep):
    """recover_from_repr_smpl(recover_mode='joint_rel_traj'), up_axis='z' (motion_representation.py:308-329, 355-371)."""
    rot_vel = rep['root_rot_angle_vel'][..., 0]
    ang = torch.zeros_like(rot_vel)
    ang[..., 1:] = rot_vel[..., :-1]
    ang = torch.cumsum(ang, dim=-1)
    q = torch.zeros(ang.shape + (4,), dtype=ang.dtype)
    q[..., 0] = torch.cos(ang)
    q[..., 3] = torch.sin(ang)
    r_pos = torch.zeros(ang.shape + (3,), dtype=ang.dtype)
    r_pos[..., 1:, [0, 1]] = rep['root_l_vel'][..., :-1, :]
    r_pos = ko.qrot(ko.qinv(q), r_pos)
    r_pos = torch.cumsum(r_pos, dim=-2)
    r_pos[..., 2] = rep['root_height'][..., 0]
    pos = rep['local_positions'][..., 3:]
    pos = pos.reshape(pos.shape[:-1] + (21, 3))
    pos = ko.qrot(ko.qinv(q[..., None, :]).expand(pos.shape[:-1] + (4,)), pos)
    pos = torch.cat([pos[..., 0:1] + r_pos[..., None, 0:1], pos[..., 1:2] + r_pos[..., None, 1:2], pos[..., 2:3]], dim=-1)
    return torch.cat([r_pos.unsqueeze(-2), pos], dim=-2)


# ------------------------------------------------------------------------------------------------------------
# 2-D reprojection guidance
# ------------------------------------------------------------------------------------------------------------
PROJ_JOINTS = [16, 18, 20, 17, 19, 21, 4, 5, 7, 8]


def guide_projection(x0, mean, std, model, transf_matrix, cam_R, cam_t, focal, center, keypoints_2d, traj_feat_dim=22):
    """grad of -(2-D loss) w.r.t. the normalised x0 [B,294,1,T] (posenet.py:260-317), plus the loss value."""
    x = x0.detach().clone().requires_grad_()
    full = x[:, :, 0].permute(0, 2, 1) * std + mean
    rep = ko.split_repr(full)
    j = ko.joints_from_smplx(rep, model)  # [B,T,22,3]
    B, T = j.shape[0], j.shape[1]
    c2s = torch.linalg.inv(transf_matrix)
    R, t = c2s[:, 0:3, 0:3], c2s[:, 0:3, -1]
    js = torch.einsum('bij,btkj->btki', R, j) + t[:, None, None, :]
    jc = torch.einsum('ij,btkj->btki', torch.linalg.inv(cam_R), js - cam_t.reshape(1, 1, 1, 3))
    pr = jc / jc[..., -1:]
    u = focal[:, None, None, 0] * pr[..., 0] + center[:, None, None, 0]
    v = focal[:, None, None, 1] * pr[..., 1] + center[:, None, None, 1]
    p2d = torch.stack([u, v], dim=-1)
    loss = (p2d - keypoints_2d[:, :T, :, 0:2]).abs() * keypoints_2d[:, :T, :, [-1]]
    loss = loss[:, :, PROJ_JOINTS].mean()
    g = torch.autograd.grad([-loss], [x])[0]
    g[:, 0:traj_feat_dim] = 0
    g[:, -4:] = 0
    return g, loss.detach()
