"""Belief-state model functions on the device (rows H1/H2): thin wrappers over the bmpc_hmm_* entry points."""
import ctypes as C

import numpy as np

from . import abi


def _dev(a, device):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=torch.device("cuda", device))


def _kinds(kinds):
    return (C.c_int32 * len(kinds))(*[int(k) for k in kinds])


def backup_rollout(x0, kinds, N, dt, Kpsi, device=0):
    """x0 (count, M, 4) -> xbackup (count, M*m, N*4) (HMM_backup_dyn.PredictiveModel.generate_backup_traj)."""
    import torch
    lib = abi.load_library()
    x0 = np.asarray(x0, dtype=np.float64)
    count, M, m = x0.shape[0], x0.shape[1], len(kinds)
    tx = _dev(x0, device)
    out = torch.empty((count, M * m, N * 4), dtype=torch.float64, device=tx.device)
    rc = lib.bmpc_hmm_backup_rollout(tx.data_ptr(), count, M, m, _kinds(kinds), N, dt, Kpsi, out.data_ptr(), device, None)
    if rc != abi.OK:
        raise RuntimeError("bmpc_hmm_backup_rollout failed (%d)" % rc)
    return out.cpu().numpy()


def rollout_sensitivity(x0, kinds, steps, ts, f0, Kpsi, device=0):
    import torch
    lib = abi.load_library()
    x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
    count, m = x0.shape[0], len(kinds)
    tx, tf = _dev(x0, device), _dev(f0, device)
    xx = torch.empty((count, m, steps, 4), dtype=torch.float64, device=tx.device)
    QQ = torch.empty((count, m, steps, 4, 4), dtype=torch.float64, device=tx.device)
    Qt = torch.empty((count, m, steps, 4), dtype=torch.float64, device=tx.device)
    rc = lib.bmpc_hmm_rollout_sensitivity(tx.data_ptr(), count, m, _kinds(kinds), steps, ts, Kpsi, tf.data_ptr(),
                                          xx.data_ptr(), QQ.data_ptr(), Qt.data_ptr(), device, None)
    if rc != abi.OK:
        raise RuntimeError("bmpc_hmm_rollout_sensitivity failed (%d)" % rc)
    return xx.cpu().numpy(), QQ.cpu().numpy(), Qt.cpu().numpy()


def belief_update(ego, xb, b, params, cbf=None, clip=True, device=0):
    """ego (count,4), xb (count,M,m,4), b (count,M,m) -> h, H, b_next."""
    import torch
    lib = abi.load_library()
    b = np.asarray(b, dtype=np.float64)
    count, M, m = b.shape
    te, tb, tbel = _dev(ego, device), _dev(xb, device), _dev(b, device)
    tc = None if cbf is None else _dev(cbf, device)
    h = torch.empty((count, M, m), dtype=torch.float64, device=te.device)
    H = torch.empty((count, M, m, m), dtype=torch.float64, device=te.device)
    bn = torch.empty((count, M, m), dtype=torch.float64, device=te.device)
    p = abi.HmmParams(**params)
    rc = lib.bmpc_hmm_belief_update(te.data_ptr(), tb.data_ptr(), tbel.data_ptr(), None if tc is None else tc.data_ptr(), count,
                                    M, m, C.byref(p), int(bool(clip)), h.data_ptr(), H.data_ptr(), bn.data_ptr(), device, None)
    if rc != abi.OK:
        raise RuntimeError("bmpc_hmm_belief_update failed (%d)" % rc)
    torch.cuda.synchronize(te.device)
    return h.cpu().numpy(), H.cpu().numpy(), bn.cpu().numpy()
