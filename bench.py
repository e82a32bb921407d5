#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200 (BASELINE.json): UTF-8 input GB/s (+ tokens/s) of batch encode.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mb MB] [--workload bpe|unigram|pipeline]

One step = one pass of aksharTokenizer.encode over one batch: normalize_text -> BPE-24k ids for every row of the
batch (BASELINE.json configs[1]: "BPE vocab 24k batch encode of 1 GB synthetic Hinglish, 1 B200").  With N > 1 every
rank (one process per GPU, launched by torchrun) encodes its own 1 GB shard of sentences -- weak scaling, no collective
on the data path; only the timing is reduced (max over ranks).

Prints ONE JSON line.  `value` is device-timed with the input already in HBM; `e2e` is the same metric through the
public batch API from pinned host buffers with the H2D copy of the text and the D2H read of the ids inside the
timed region.  `roofline` is the dominant kernel's algorithmic bytes / its CUDA-event time against
MEASURED_PEAKS.json; `cpu_baseline` is the oracle port of the same path on the host cores (bounded sample).
--impl reference times that CPU implementation alone, with every host core.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'tools')):
    if p not in sys.path:
        sys.path.insert(0, p)

MODELS = os.path.join(ROOT, 'tests', 'golden', 'models')
# DRAM traffic per INPUT byte of each hot kernel: (dram__bytes_read.sum + dram__bytes_write.sum) / input bytes from the
# `ncu --set full` captures at 256 MiB summarised in profiles/r01_ncu_full_summary.csv; scaled to the run's size below
TRAFFIC_PER_INPUT_BYTE = {
    'ak_nf3_classify_kernel': (292.68 + 63.60) / 268.44,
    'ak_nf_write_kernel': (372.49 + 248.46) / 268.44,
    'ak_bf3_encode_kernel': (446.12 + 398.04) / 268.44,
    'ak_sf3_kernel': (301.27 + 496.13) / 268.44,
}
CHUNK = 32 << 20
SEED = 20261018


# ------------------------------------------------------------------ synthetic corpus
_corpus = {}


def _gen_chunk(args):
    kind, index, nbytes = args
    import synth_corpus as sc
    if kind not in _corpus:
        _corpus[kind] = sc.Corpus(kind, SEED)
    return _corpus[kind].chunk(index, nbytes)


def make_corpus(kind, nbytes, first_chunk=0, procs=None):
    """-> (uint8 numpy array, int64 row offsets); chunk indices first_chunk.. (rank-disjoint for the sharded runs)"""
    import numpy as np
    n = max(1, (nbytes + CHUNK - 1) // CHUNK)
    jobs = [(kind, first_chunk + i, min(CHUNK, nbytes - i * CHUNK)) for i in range(n)]
    procs = procs or min(len(jobs), os.cpu_count() or 1)
    if procs > 1:
        with mp.get_context('fork').Pool(procs) as pool:
            parts = pool.map(_gen_chunk, jobs)
    else:
        parts = [_gen_chunk(j) for j in jobs]
    datas, offs, base = [], [np.zeros(1, dtype=np.int64)], 0
    for d, o in parts:
        datas.append(d)
        offs.append(o[1:] + base)
        base += d.size
    return np.concatenate(datas), np.concatenate(offs)


# ------------------------------------------------------------------ CPU arm: the oracle port on the host cores
def _cpu_init(workload):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import akshar_oracle as O
    global _O, _M
    _O = O
    if workload == 'bpe':
        _M = O.BpeModel(os.path.join(MODELS, 'bpe24k.json'))
    elif workload == 'unigram':
        _M = O.UnigramModel(os.path.join(MODELS, 'spm24k.model'))
    else:
        _M = None


def _cpu_work(args):
    workload, lines = args
    n = 0
    for s in lines:
        norm = _O.normalize_text(s)
        if workload == 'bpe':
            n += len(_O.bpe_encode(_M, norm))
        elif workload == 'unigram':
            n += len(_O.unigram_encode(_M, norm))
        else:
            cps = [ord(c) for c in norm]
            n += len(_O.grapheme_breaks(cps)) + len(_O.script_runs(cps))
    return n


class CpuArm:
    """the oracle restatement of the same path, one process per host core, over a bounded sample of the workload"""

    def __init__(self, workload, data, off, sample_bytes):
        import numpy as np
        self.workload = workload
        hi = int(np.searchsorted(off, sample_bytes, side='right'))
        hi = max(1, min(hi, off.size - 1))
        b = data[:off[hi]].tobytes()
        self.lines = [b[off[i]:off[i + 1]].decode('utf-8') for i in range(hi)]
        self.nbytes = int(off[hi])
        self.cores = os.cpu_count() or 1
        self.pool = mp.get_context('fork').Pool(self.cores, initializer=_cpu_init, initargs=(workload,))
        per = max(1, len(self.lines) // (self.cores * 4))
        self.jobs = [(workload, self.lines[i:i + per]) for i in range(0, len(self.lines), per)]

    def step(self):
        t = time.perf_counter()
        tokens = sum(self.pool.map(_cpu_work, self.jobs))
        return time.perf_counter() - t, tokens

    def close(self):
        self.pool.close()
        self.pool.join()


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows = []             # (arrival time, line)
        self.proc = None
        self.t_from = 0.0
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=5.0):
        """nvidia-smi takes a few hundred ms to deliver its first line: wait for it so that the (short) measured phases
        are covered"""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        """only samples that arrive from now on count (the measured phases start here)"""
        self.t_from = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, r in self.rows:
            if ts < self.t_from:
                continue
            f = [x.strip() for x in r.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


# ------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mb', type=int, default=1024, help='synthetic input per GPU in MiB (1024 = the configuration the metric is quoted on)')
    ap.add_argument('--workload', default='bpe', choices=['bpe', 'unigram', 'pipeline'])
    ap.add_argument('--cpu-sample-mb', type=float, default=24.0)
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    kind = {'bpe': 'hinglish', 'unigram': 'hindi', 'pipeline': 'social'}[a.workload]
    cfg = {'workload': {'bpe': 'BPE-24k batch encode (normalize_text -> ids) of %d MiB synthetic Hinglish per GPU' % a.mb,
                        'unigram': 'Unigram-24k Viterbi encode of %d MiB synthetic Hindi per GPU' % a.mb,
                        'pipeline': 'normalize + akshar + code-switch on %d MiB synthetic social Hinglish per GPU' % a.mb}[a.workload],
           'bytes_per_gpu': a.mb << 20, 'model': {'bpe': 'tests/golden/models/bpe24k.json', 'unigram': 'tests/golden/models/spm24k.model',
                                                  'pipeline': None}[a.workload],
           'l2': 'inputs larger than L2 (no flush needed)' if a.mb >= 256 else 'input smaller than 2x L2', 'sharding': 'sentences, no collective'}
    metric = 'utf8_input_GBps_batch_encode' if a.workload != 'pipeline' else 'utf8_input_GBps_normalize_segment'

    if a.impl == 'reference':
        if rank != 0:
            return
        sample = int(a.cpu_sample_mb * (1 << 20))
        data, off = make_corpus(kind, sample + (1 << 20), 0)
        arm = CpuArm(a.workload, data, off, sample)
        for _ in range(a.warmup):
            arm.step()
        t = tok = 0.0
        for _ in range(a.steps):
            dt, n = arm.step()
            t += dt
            tok += n
        arm.close()
        v = arm.nbytes * a.steps / t / 1e9
        print(json.dumps({
            'impl': 'reference', 'metric': metric, 'value': v, 'unit': 'GB/s', 'tokens_per_s': tok / t, 'n_gpus': a.gpus,
            'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': t / a.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic', 'config': cfg,
            'cpu_baseline': {'value': v, 'unit': 'GB/s', 'cores': arm.cores, 'kind': 'port',
                             'sample': '%d rows / %.1f MB of the same synthetic workload per step; oracle/akshar_oracle.py '
                                       '(pure-Python restatement; the reference itself is Python and cannot travel to this box)'
                                       % (len(arm.lines), arm.nbytes / 1e6)},
            'e2e': {'value': v, 'unit': 'GB/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    # host-side work that forks (corpus generation, the CPU arm's worker pool) happens before CUDA is touched
    nbytes = a.mb << 20
    chunks_per_rank = (nbytes + CHUNK - 1) // CHUNK
    data, off = make_corpus(kind, nbytes, rank * chunks_per_rank, procs=max(1, (os.cpu_count() or 1) // world))
    n_rows = off.size - 1
    arm = CpuArm(a.workload, data, off, int(a.cpu_sample_mb * (1 << 20))) if rank == 0 else None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    import akshar_b200 as A
    from akshar_b200 import _lib as C

    h_data = torch.from_numpy(data).pin_memory()
    h_off = torch.from_numpy(off).pin_memory()
    if a.workload == 'bpe':
        tk = A.aksharTokenizer(os.path.join(MODELS, 'bpe24k.json'), 'bpe', device=local)
        eng, mkind = tk._eng, 0
    elif a.workload == 'unigram':
        tk = A.aksharTokenizer(os.path.join(MODELS, 'spm24k.model'), 'sentencepiece', device=local)
        eng, mkind = tk._eng, 1
    else:
        eng, mkind = A.Engine(local), None
    dev_batch = eng.put((h_data, h_off))
    torch.cuda.synchronize()

    def device_step():
        if mkind is not None:
            return eng.tokenizer_encode_batch(dev_batch, mkind, check=False)
        norm, r1 = eng.normalize_batch(dev_batch, check=False)
        # the normalized length stays on the device in the fused entry point; here (two ABI calls) it is read back once
        total = int(r1[0].item())
        norm.end = total
        c, r, r2 = eng.segment_batch(norm, clusters=True, runs=True, check=False)
        return (c, r), norm, r2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness of the run that is about to be timed: status bits clear, totals sane
    out = device_step()
    torch.cuda.synchronize()
    res = out[-1].cpu()
    assert int(res[2]) == 0, 'status bits %d' % int(res[2])
    n_tokens = int(res[0]) + (int(res[1]) if mkind is None else 0)
    n_norm = int(res[1]) if mkind is not None else int(out[1].end)
    del out
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.wait_first()
    for _ in range(max(0, a.warmup - 1)):
        device_step()
    barrier()
    if sampler:
        sampler.mark()          # clocks are sampled every 20 ms from here to the end of the measured phases (timed steps,
    l0 = eng.launch_count()     # end-to-end steps, per-kernel timing steps): the timed region alone lasts < 100 ms
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        device_step()
    e1.record()
    barrier()
    launches = eng.launch_count() - l0
    ms = e0.elapsed_time(e1) / a.steps
    # ---- end to end through the public batch API: pinned host text in, ids (or offsets) on the host out
    def e2e_step():
        if mkind is not None:
            # the public host -> ids call: H2D of chunk k+1, kernels of chunk k and D2H of chunk k-1 overlap
            ids, splits = eng.encode_host_pipelined(h_data, h_off, mkind)
            assert int(splits[-1]) == ids.numel()
            return ids.numel() * 4 + splits.numel() * 8 + 32
        b = eng.put((h_data, h_off))
        norm, r1 = eng.normalize_batch(b, check=False)
        norm.end = int(r1[0].item())
        c, r, r2 = eng.segment_batch(norm, clusters=True, runs=True, check=False)
        rr = r2.cpu()
        nc, nr = int(rr[0]), int(rr[1])
        if not hasattr(e2e_step, 'bufs'):
            e2e_step.bufs = [torch.empty(nc, dtype=torch.int32).pin_memory(), torch.empty(nr, dtype=torch.int32).pin_memory(),
                             torch.empty(nr, dtype=torch.uint8).pin_memory(), torch.empty(n_rows + 1, dtype=torch.int64).pin_memory(),
                             torch.empty(n_rows + 1, dtype=torch.int64).pin_memory(), torch.empty(norm.end, dtype=torch.uint8).pin_memory()]
        hb = e2e_step.bufs
        hb[0][:nc].copy_(c.values[:nc], non_blocking=True)
        hb[1][:nr].copy_(r.values[:nr], non_blocking=True)
        hb[2][:nr].copy_(r.extra[:nr], non_blocking=True)
        hb[3].copy_(c.splits, non_blocking=True)
        hb[4].copy_(r.splits, non_blocking=True)
        hb[5][:norm.end].copy_(norm.data[:norm.end], non_blocking=True)
        torch.cuda.synchronize()
        return nc * 4 + nr * 5 + 2 * (n_rows + 1) * 8 + norm.end + 64

    d2h = e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(a.steps, 3))
    for _ in range(e2e_steps):
        d2h = e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps

    # ---- dominant kernel: the library brackets its hot kernels with CUDA events on the launching stream
    # (akshar_timing_enable); average over a few full steps of the same workload
    eng.timing(True)
    names = ['ak_nf3_classify_kernel', 'ak_nf_write_kernel'] + (['ak_words_kernel', 'ak_resolve_kernel<bpe>', 'ak_emit_kernel'] if mkind == 0 else
                                                                 ['ak_words_kernel', 'ak_resolve_kernel<unigram>', 'ak_emit_kernel'] if mkind == 1 else ['ak_sf3_kernel'])
    acc = {k: [] for k in names}
    for _ in range(3):
        device_step()
        torch.cuda.synchronize()
        for k in names:
            v = eng.kernel_ms(k)
            if v is not None:
                acc[k].append(v)
    eng.timing(False)
    clocks = sampler.stop() if sampler else None
    kms = {k: sum(v) / len(v) for k, v in acc.items() if v}
    n_c = int(res[0]) if mkind is None else 0
    n_r = int(res[1]) if mkind is None else 0
    # algorithmic bytes per launch: logical input read once + required output written once
    alg = {
        'ak_nf3_classify_kernel': nbytes + 4 * (nbytes // 15),                 # text in, one 4-byte emit mask per 16-byte chunk out
        'ak_nf_write_kernel': nbytes + 4 * (nbytes // 15) + n_norm + 8 * (n_rows + 1),
        'ak_words_kernel': n_norm + 8 * (n_rows + 1),
        'ak_resolve_kernel<bpe>': n_norm + 4 * n_tokens + 8 * (n_rows + 1),
        'ak_resolve_kernel<unigram>': n_norm + 4 * n_tokens + 8 * (n_rows + 1),
        'ak_emit_kernel': 4 * n_tokens + 8 * (n_rows + 1),
        'ak_sf3_kernel': n_norm + 4 * n_c + 5 * n_r + 16 * (n_rows + 1),
    }
    stages = {k: (kms[k], alg[k]) for k in kms}
    dom = max(stages, key=lambda k: stages[k][0])
    peak, peak_kind = peaks()
    ach = stages[dom][1] / (stages[dom][0] * 1e-3) / 1e9

    # ---- reduce over ranks
    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device='cuda')
    cnt = torch.tensor([nbytes, n_tokens, launches, d2h], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms = float(t[0]), float(t[1])
    tot_bytes, tot_tokens, tot_launch, tot_d2h = (float(x) for x in cnt)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    arm.step()
    dt, ctok = arm.step()
    arm.close()
    cpu = {'value': arm.nbytes / dt / 1e9, 'unit': 'GB/s', 'cores': arm.cores, 'kind': 'port', 'tokens_per_s': ctok / dt,
           'sample': '%d rows / %.1f MB of this workload, oracle/akshar_oracle.py on %d processes, %.1f s'
                     % (len(arm.lines), arm.nbytes / 1e6, arm.cores, dt)}
    line = {
        'metric': metric, 'value': tot_bytes / (ms_max * 1e-3) / 1e9, 'unit': 'GB/s', 'tokens_per_s': tot_tokens / (ms_max * 1e-3),
        'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms_max, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic', 'config': cfg,
        'rows_per_gpu': n_rows, 'tokens_per_gpu': n_tokens, 'normalized_bytes_per_gpu': n_norm,
        'e2e': {'value': tot_bytes / (e2e_ms * 1e-3) / 1e9, 'unit': 'GB/s', 'h2d_bytes_per_step': int(nbytes + 8 * (n_rows + 1)),
                'd2h_bytes_per_step': int(tot_d2h / world), 'ms_per_step': e2e_ms},
        'gpu_launches': int(tot_launch),
        'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': ach, 'peak': peak, 'peak_source': peak_kind, 'unit': 'GB/s',
                     'frac': ach / peak, 'frac_of_nominal_8000': ach / 8000.0,
                     'traffic': (int(TRAFFIC_PER_INPUT_BYTE[dom] * nbytes) if dom in TRAFFIC_PER_INPUT_BYTE else None),
                     'traffic_source': 'ncu --set full at 256 MiB, scaled by input bytes (profiles/r01_v3_ncu_full_summary.csv)',
                     'ms': stages[dom][0], 'algorithmic_bytes': stages[dom][1],
                     'kernels_ms': {k: v[0] for k, v in stages.items()},
                     'kernels_frac': {k: v[1] / (v[0] * 1e-3) / 1e9 / peak for k, v in stages.items()}},
        'cpu_baseline': cpu, 'clocks': clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
