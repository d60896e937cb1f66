"""Helpers shared by the -m gpu parity tests (CUDA path through the C ABI vs the oracle)."""
import numpy as np
import torch

from surrdamh_b200 import _lib as L
from surrdamh_b200.chains import ChainBlock, Proposal, Trace, compact_snapshots
from surrdamh_b200.device import DeviceModel, DeviceProblem


def problem_from_kw(d, kw):
    return DeviceProblem(d, prior_mean=kw["prior_mean"], prior_std=kw["prior_std"], noise_std=kw["noise_std"],
                         no_observations=kw["no_observations"], observations=kw["observations"])


def pack_streams(per_chain):
    """[(z [K x d], u [K x 2]) per chain] -> device z [K][d][C], uni [K][2][C]."""
    z = np.stack([p[0] for p in per_chain], axis=-1)       # [K][d][C]
    u = np.stack([p[1] for p in per_chain], axis=-1)       # [K][2][C]
    return L.dev(np.ascontiguousarray(z)), L.dev(np.ascontiguousarray(u))


def run_block(kind, d, m, kw, truth, surrogate, K, std, per_chain_streams, init, snapshots=True, lanes=0,
              per_chain_std=False, split=None):
    """Run C chains for K steps under explicit streams; returns (block, trace).
    split: None -> fused device G; otherwise a callable params [n x d] -> G [n x m] evaluated on
    the host for the stage-1 survivors (the reference's solver plugin path)."""
    C_ = len(per_chain_streams)
    prob = problem_from_kw(d, kw)
    blk = ChainBlock(prob, C_, initial=init)
    prop = Proposal(d, std, per_chain=per_chain_std)
    z, uni = pack_streams(per_chain_streams)
    tr = Trace(K, C_, d, m, blk.device, snapshots=snapshots)
    if split is None:
        blk.prepare(kind, surrogate, truth)
        blk.run(kind, K, prop, surrogate, truth, trace=tr, streams=(z, uni), lanes=lanes)
        torch.cuda.synchronize()
        return blk, tr
    # split path: one step per launch pair, G from the host callable
    blk.prepare(kind, surrogate, None, G_host=split(blk.current_numpy()))
    one = Trace(1, C_, d, m, blk.device, snapshots=snapshots)
    for k in range(K):
        st = (z[k:k + 1].contiguous(), uni[k:k + 1].contiguous())
        blk.propose(kind, prop, surrogate, one, streams=st)
        flags = one.flags[0].cpu().numpy()
        props = one.prop[0].cpu().numpy().T            # [C][d]
        G = np.zeros((C_, m))
        idx = np.nonzero(flags == L.FLAG_PENDING)[0]
        if idx.size:
            G[idx] = split(props[idx])
        Gd = L.dev(np.ascontiguousarray(G.T))
        blk.decide(kind, prop, surrogate, one, Gd, streams=st)
        tr.flags[k] = one.flags[0]
        tr.prop[k], tr.post[k], tr.pre[k] = one.prop[0], one.post[0], one.pre[0]
        if snapshots:
            tr.snap_par[k], tr.snap_obs[k] = one.snap_par[0], one.snap_obs[0]
    torch.cuda.synchronize()
    return blk, tr


def chain_view(tr, c):
    """(flags [K], props [K x d], post [K], pre [K]) of chain c."""
    return (tr.flags[:, c].cpu().numpy(), tr.prop[:, :, c].cpu().numpy(), tr.post[:, c].cpu().numpy(),
            tr.pre[:, c].cpu().numpy())


def chain_snapshots(tr, c):
    f = tr.flags[:, c].cpu().numpy()
    keep = f != L.FLAG_PREREJECTED
    return tr.snap_par[:, :, c].cpu().numpy()[keep], tr.snap_obs[:, :, c].cpu().numpy()[keep], f[keep]


def data_rows(tr, blk_initial, post0, c):
    """Rebuild the compressed-chain rows [weight, u..., post, pre_post] of chain c from the trace
    (what __write_to_file__ emits at every acceptance and at the end of the stage)."""
    flags, props, post, pre = chain_view(tr, c)
    rows, cur, cur_post, w = [], np.array(blk_initial, dtype=float), post0, 1
    for k in range(len(flags)):
        if flags[k] == L.FLAG_ACCEPTED:
            rows.append([w] + list(cur) + [cur_post])
            cur, cur_post, w = props[k], post[k], 1
        else:
            w += 1
    rows.append([w] + list(cur) + [cur_post])
    return np.array(rows)
