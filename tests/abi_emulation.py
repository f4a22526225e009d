"""CPU test doubles that follow the *semantics* of the C ABI (include/csa_b200.h) with plain torch, so that host-side
logic that strings several native calls together (spider_b200/dist.py) can be exercised without a GPU.  Test
infrastructure only: nothing under spider_b200/ imports this, and the GPU parity tests never use it."""
import torch

from oracle import reference_port as rp

TILE = 128


def idx_stride_for(n_cols):
    return (n_cols + TILE - 1) // TILE * TILE


def compact_rows(mask_rows, n_rows, n_cols, row_stride, block_n=0, limit_cols=0, idx=None, counts=None):
    flat = mask_rows.reshape(-1).to(torch.uint8)
    stride = idx_stride_for(n_cols)
    if idx is None:      # like the ABI: the caller's buffers are filled in place when it brings them
        idx = torch.zeros((n_rows, stride), dtype=torch.int32)
    if counts is None:
        counts = torch.zeros((n_rows,), dtype=torch.int32)
    for r in range(n_rows):
        row = flat[r * row_stride:r * row_stride + n_cols].bool().clone()
        if block_n > 0:
            row[limit_cols:] = False
            row[r * block_n:(r + 1) * block_n] = True
        nz = torch.nonzero(row)[:, 0].int()
        idx[r, :nz.numel()] = nz
        counts[r] = nz.numel()
    return idx, counts


def sample_ranges(s_idx, s_count, block_n, n_frames, out=None):
    c = int(s_count.item())
    S = s_idx.reshape(-1)[:c].long()
    out = torch.zeros((n_frames + 1, 4), dtype=torch.int32) if out is None else out.view(n_frames + 1, 4)
    for f in range(n_frames):
        lo = int(torch.searchsorted(S, torch.tensor(f * block_n)))
        hi = int(torch.searchsorted(S, torch.tensor((f + 1) * block_n)))
        out[f] = torch.tensor([0, lo, hi, c - hi], dtype=torch.int32)
    out[n_frames] = torch.tensor([0, c, 0, 0], dtype=torch.int32)
    return out


def gather_rows(src, idx, max_rows, row_base=0, count=None, count_adjust=0, out=None):
    n = max_rows if count is None else min(max_rows, int(count.item()) + count_adjust)
    if out is None:
        out = torch.empty((max_rows, src.shape[1]), dtype=src.dtype)
    out[:n] = src[row_base + idx[:n].long()]
    return out


def gather_kv(k, v, group_rows, n_groups, s_idx, s_count, max_rows):
    c = min(int(s_count.item()), max_rows)
    cap = max_rows + TILE
    s_idx = s_idx.reshape(-1)
    # rows beyond the zero tail are "undefined" in the ABI: poison them so that a reader of them is caught
    k_s = torch.full((n_groups * cap, k.shape[1]), float("nan"), dtype=k.dtype)
    v_s = torch.full((n_groups * cap, k.shape[1]), float("nan"), dtype=k.dtype)
    for g in range(n_groups):
        k_s[g * cap:g * cap + c] = k[g * group_rows + s_idx[:c].long()]
        v_s[g * cap:g * cap + c] = v[g * group_rows + s_idx[:c].long()]
        k_s[g * cap + c:min((g + 1) * cap, g * cap + c + TILE)] = 0
        v_s[g * cap + c:min((g + 1) * cap, g * cap + c + TILE)] = 0
    return k_s, v_s, cap


def attn_fwd(q, o, *, heads, n_groups, n_frames, n_q, k_a=None, v_a=None, a_group_rows=0, k_b=None, v_b=None,
             b_group_rows=0, idx=None, counts=None, list_base=-1, list_step=0, g_adjust=0, ca=(0, 0, 0),
             cb=(0, 0, 0), scale=None, max_ctas=0, ranges=None, range_base=0, range_step=0, split=True,
             b_first=False, ready=None, ready_epoch=0, ready_bounds=None, ready_peers=0, ready_frames_per_peer=0,
             epoch_base=None, done=None, done_counter=None, peer_self=0):
    C = q.shape[1]
    if epoch_base is not None:
        ready_epoch += int(epoch_base[0])
    if ready is not None:
        # the kernel waits per peer when it reaches that peer's rows; the emulation waits for all of them up front
        for r in range(ready_peers if ready_frames_per_peer > 0 else len(ready_bounds) - 1):
            _wait_ge(ready, r, ready_epoch, "ready")
    for g in range(n_groups):
        for f in range(n_frames):
            ks, vs = [], []
            if list_base >= 0:
                li = list_base + f * list_step
                n = max(0, int(counts[li]) + g_adjust)
                rows = g * a_group_rows + idx[li, :n].long()
                ks.append(k_a[rows]); vs.append(v_a[rows])
            if ranges is not None:
                r = ranges[range_base + f * range_step].tolist()
                for st, ln in ((r[0], r[1]), (r[2], r[3])):
                    if ln > 0:
                        ks.append(k_a[g * a_group_rows + st:g * a_group_rows + st + ln])
                        vs.append(v_a[g * a_group_rows + st:g * a_group_rows + st + ln])
            elif ca[2] > 0:
                st = g * a_group_rows + ca[0] + f * ca[1]
                ks.append(k_a[st:st + ca[2]]); vs.append(v_a[st:st + ca[2]])
            if cb[2] > 0:
                st = g * b_group_rows + cb[0] + f * cb[1]
                ks.append(k_b[st:st + cb[2]]); vs.append(v_b[st:st + cb[2]])
            kk, vv = torch.cat(ks).float(), torch.cat(vs).float()
            assert torch.isfinite(kk).all() and torch.isfinite(vv).all(), "attention read an undefined K/V row"
            qs = slice((g * n_frames + f) * n_q, (g * n_frames + f + 1) * n_q)
            keys = [torch.arange(kk.shape[0], dtype=torch.int32)]
            out = rp.gathered_attention(q[qs].float()[None], kk[None], vv[None], keys, heads)[0]
            o[qs] = out.to(o.dtype)
    if done_counter is not None:
        # the launch's last CTA releases the exchange buffers of this epoch on every peer (csa_attn_args_t.done_dst)
        assert ready is not None and int(done_counter[0]) == 0
        for r in range(len(done)):
            if r != peer_self:
                done[r][peer_self] = ready_epoch
        PEER_LOG.append(("release", ready_epoch))
    return o


def gemm(x, w, bias=None, out=None, alpha=1.0, scatter=None, out2=None, exchange=None):
    """csa_gemm: y = alpha * x w^T (+ bias), optional second output, optional fused gather — local (scatter buffers)
    or as the multi-GPU exchange (csa_peer_exchange_t: every peer's K[S] / V[S], arrival flags, release wait)."""
    m, n = x.shape[0], w.shape[0]
    y = alpha * (x.float() @ w.float().t())
    if bias is not None:
        y = y + bias.float()
    y = y.to(x.dtype)
    if out2 is not None:
        n1 = out.shape[1]
        out.copy_(y[:, :n1])
        out2.copy_(y[:, n1:])
    else:
        if out is None:
            out = torch.empty((m, n), dtype=x.dtype)
        out.copy_(y)
    if scatter is not None:
        pos, k_s, v_s, group_rows, dst_group_rows, split_col = scatter[:6]
        col0 = scatter[6] if len(scatter) > 6 else 0
        if exchange is not None:
            epoch, done_epoch = exchange["epoch"], exchange["done_epoch"]
            eb = exchange.get("epoch_base")
            if eb is not None:
                epoch += int(eb[0])
                done_epoch += int(eb[0])
            me = exchange["self"]
            assert m == group_rows and int(exchange["counter"][0]) == 0
            for r in range(len(exchange["k_dst"])):
                if r != me and done_epoch > 0:
                    _wait_ge(exchange["done"], r, done_epoch, "done")
            sel = (pos[:m] >= 0).nonzero().flatten()
            dst = pos[:m][sel].long()
            for r in range(len(exchange["k_dst"])):
                exchange["k_dst"][r][dst] = y[sel, col0:split_col]
                exchange["v_dst"][r][dst] = y[sel, split_col:]
            for r in range(len(exchange["k_dst"])):
                exchange["ready"][r][me] = epoch
            PEER_LOG.append(("exchange", epoch, int(sel.numel()), done_epoch))
        else:
            for g in range(m // group_rows):
                pg = pos[:group_rows]
                sel = (pg >= 0).nonzero().flatten()
                dst = g * dst_group_rows + pg[sel].long()
                k_s[dst] = y[g * group_rows + sel, col0:split_col]
                v_s[dst] = y[g * group_rows + sel, split_col:]
    return out if out2 is None else (out, out2)


def sample_positions(s_idx, s_count, n_cols, out=None):
    c = min(int(s_count[0]) if s_count.dim() else int(s_count), n_cols)
    pos = torch.full((n_cols,), -1, dtype=torch.int32) if out is None else out.fill_(-1)
    pos[s_idx[:c].long()] = torch.arange(c, dtype=torch.int32)
    return pos


def _wait_ge(flags, i, value, what, timeout=120.0):
    import time
    t0 = time.time()
    while int(flags[i]) < value:
        if time.time() - t0 > timeout:
            raise TimeoutError(f"{what}[{i}] stayed at {int(flags[i])} < {value}")
        time.sleep(0.001)


def peer_scatter_kv(k, v, idx, count, dst_row0, k_dst, v_dst, ready, self_index, epoch, done, done_epoch, counter,
                    ranges=None, frames_per_peer=0, idx_adjust=0, epoch_base=None):
    if epoch_base is not None:     # epochs in device memory: offsets to the base (done_epoch may be <= 0)
        epoch += int(epoch_base[0])
        done_epoch += int(epoch_base[0])
    for r in range(len(k_dst)):
        if r != self_index and done_epoch > 0:
            _wait_ge(done, r, done_epoch, "done")
    if ranges is not None:   # device-side geometry of the ABI: this rank's run of the sampled list
        dst_row0 = int(ranges[self_index * frames_per_peer][1])
        count = int(ranges[(self_index + 1) * frames_per_peer - 1][2]) - dst_row0
        idx = idx[dst_row0:dst_row0 + count] + idx_adjust
    rows = idx[:count].long()
    for r in range(len(k_dst)):
        # poison what a correct reader never touches before the flag is up: the rows are written AFTER a delay only
        # in the sense that the flag follows them
        k_dst[r][dst_row0:dst_row0 + count] = k[rows]
        v_dst[r][dst_row0:dst_row0 + count] = v[rows]
    for r in range(len(k_dst)):
        ready[r][self_index] = epoch
    PEER_LOG.append(("scatter", epoch, count, dst_row0, done_epoch))


def epoch_advance(epoch_base, delta):
    epoch_base[0] += delta
    PEER_LOG.append(("advance", delta))


def peer_signal(done, self_index, epoch, like, epoch_base=None):
    if epoch_base is not None:
        epoch += int(epoch_base[0])
    for r in range(len(done)):
        if r != self_index:
            done[r][self_index] = epoch
    PEER_LOG.append(("signal", epoch))


def enable_peer_access(peer_device):
    pass


PEER_LOG = []


def install(monkeypatch_or_none, native, processor_cls=None):
    """Replace the native entry points by the emulations (monkeypatch fixture, or plain setattr when None)."""
    pairs = dict(compact_rows=compact_rows, sample_ranges=sample_ranges, gather_rows=gather_rows,
                 gather_kv=gather_kv, attn_fwd=attn_fwd, peer_scatter_kv=peer_scatter_kv, peer_signal=peer_signal,
                 enable_peer_access=enable_peer_access, epoch_advance=epoch_advance, gemm=gemm,
                 sample_positions=sample_positions)
    for name, fn in pairs.items():
        if monkeypatch_or_none is None:
            setattr(native, name, fn)
        else:
            monkeypatch_or_none.setattr(native, name, fn)
    if processor_cls is not None:
        processor_cls._check_input = staticmethod(lambda x: None)
