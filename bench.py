"""bench.py -- LatentAugment hot path throughput: augmented images / second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2] [--precision bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch: ``num_steps`` Adam iterations on w through the
generator (forward + backward-to-w) plus the final synthesis -- one ``LatentAugment.forward()``.
Workload at N=1: BASELINE.json configs[1] (SG2 256x256 3-ch, batch 32, 10 steps, 4096-code bank);
with N GPUs every rank runs that batch (weak scaling, no data-path collective: samples are
independent).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {          # BASELINE.json configs; oracle/synthetic.py holds the same table for the parity tests
    'c1': dict(img_resolution=128, img_channels=1, batch=4, steps=5, bank=256, img_bank=64),
    'c2': dict(img_resolution=256, img_channels=3, batch=32, steps=10, bank=4096, img_bank=64),
    'c3': dict(img_resolution=512, img_channels=3, batch=16, steps=10, bank=4096, img_bank=64),   # per-GPU shard of B=128 at 8 GPUs
    'tiny': dict(img_resolution=32, img_channels=2, batch=4, steps=3, bank=64, img_bank=8, channel_base=2048, channel_max=64),
}


def layer_table(res, img_c, channel_base=32768, channel_max=512):
    """(name, res_out, cin, cout, up) of every conv layer + algorithmic / executed MACs per image."""
    from latentaugment_b200.utils.synthetic import channels_for
    ch = channels_for(res, channel_base, channel_max)
    rows = []
    for r, c in ch.items():
        if r > 4:
            rows.append((f'b{r}.conv0', r, ch[r // 2], c, 2))
        rows.append((f'b{r}.conv1', r, c, c, 1))
    out = []
    for name, r, cin, cout, up in rows:
        if up == 2:      # algorithmic: transposed conv on the input grid + 4x4 FIR; the GEMM executes the 9 taps, the FIR is a SIMT pass
            alg = 9 * (r // 2) ** 2 * cin * cout + 16 * r * r * cout
            exe = 9 * (r // 2) ** 2 * cin * cout
        else:
            alg = exe = 9 * r * r * cin * cout
        out.append(dict(name=name, res=r, cin=cin, cout=cout, up=up, alg_macs=alg, exe_macs=exe))
    rgb = sum(r * r * c * img_c for r, c in ch.items())
    return out, rgb


def f_syn(res, img_c, **kw):
    """Forward FLOPs of one synthesis pass per image (SURVEY.md §8d definition)."""
    rows, rgb = layer_table(res, img_c, **kw)
    return 2.0 * (sum(r['alg_macs'] for r in rows) + rgb)


class ClockSampler:
    """nvidia-smi SM clock + throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.proc = None
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(',')]
            try:
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith('active'):
                        self.reasons.add(n)
            except (ValueError, IndexError):
                pass

    def mark(self):
        return len(self.samples)

    def stop(self, start=0):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        s = sorted(self.samples[start:]) or sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


def cpu_baseline(cfg_name, seconds_hint=20.0, repeat=1):
    """Reference CPU path (oracle port, fused grouped-conv formulation, torch CPU ops on all host
    threads) on a BOUNDED sample of the workload: batch 2, 2 Adam steps + final synthesis; scaled to
    the full step count by the pass count (3 generator passes per Adam step, as the reference executes:
    fprop + dgrad + per-sample wgrad; +1 final)."""
    import random

    import torch

    from oracle import latent_aug as ola
    from oracle import synthetic
    c = dict(CONFIGS[cfg_name])
    c.setdefault('channel_base', 32768)
    c.setdefault('channel_max', 512)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    bs, ss = 2, 2
    wl = synthetic.make_workload(c, noise_strength=0.0, batch=bs)
    orc = ola.LatentAugOracle(wl['G'], wl['W'], wl['X'], num_epochs=ss, fused=True)
    times = []
    random.seed(0)
    orc.forward(wl['w0'].clone())        # untimed warm-up (thread pool, primitive caches)
    for _ in range(repeat):
        random.seed(0)
        t0 = time.perf_counter()
        orc.forward(wl['w0'].clone())
        times.append(time.perf_counter() - t0)
    t = min(times)
    scale = (3 * c['steps'] + 1) / (3 * ss + 1)
    ips = bs / (t * scale)
    return {'value': ips, 'unit': 'img/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'oracle (torch CPU, fused modconv) batch {bs}, {ss} Adam steps + final synthesis in {t:.2f} s; '
                      f'scaled x{scale:.2f} to {c["steps"]} steps by generator-pass count'}, t


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    c = CONFIGS[args.config]
    vals = []
    for _ in range(args.warmup):
        cpu_baseline(args.config)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        cb, _t = cpu_baseline(args.config)
        vals.append(cb['value'])
    wall = time.perf_counter() - t_all
    v = sum(vals) / len(vals)
    cb['value'] = v
    line = {'impl': 'reference', 'metric': 'augmented images/sec', 'value': v, 'unit': 'img/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * wall / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args.config, c)}, 'cpu_baseline': cb,
            'e2e': {'value': v, 'unit': 'img/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line))


def workload_name(name, c):
    return (f'{name}: StyleGAN2 {c["img_resolution"]}x{c["img_resolution"]} {c["img_channels"]}-ch, batch {c["batch"]}/GPU, '
            f'{c["steps"]} w-opt steps + final synthesis, {c["bank"]}-code bank, {c["img_bank"]}-image bank, w_latent=w_pix=1')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c2', choices=sorted(CONFIGS))
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32_parity'])
    ap.add_argument('--w-disc', type=float, default=0.0, help='weight of the discriminator realism term (0 = the headline workload)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile', action='store_true', help='short run for ncu: skips the e2e leg, the CPU baseline and the per-GEMM timing')
    ap.add_argument('--layers-out', default='', help='write the per-layer tap-GEMM timing table (JSON) here')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from latentaugment_b200.augments import create_augment
    from latentaugment_b200.options.aug_options import AugOptions

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local_rank)
    dev = torch.device(f'cuda:{local_rank}')
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    c = dict(CONFIGS[args.config])
    B, steps, res, C = c['batch'], c['steps'], c['img_resolution'], c['img_channels']

    # ---- the reference-facing plugin, synthetic mode (random-init generator, synthetic banks)
    argv = ['--aug', 'latent', '--synthetic', '--batch_size', str(B), '--gpu_ids', str(local_rank), '--gpu_ids_aug', str(local_rank),
            '--img_resolution', str(res), '--synthetic_channels', str(C), '--synthetic_bank', str(c['bank']),
            '--synthetic_img_bank', str(c['img_bank']), '--synthetic_codes', str(max(4 * B, 256)), '--precision', args.precision,
            '--opt_num_epochs', str(steps), '--no_log',
            '--synthetic_channel_base', str(c.get('channel_base', 32768)), '--synthetic_channel_max', str(c.get('channel_max', 512))]
    # stdout carries exactly ONE JSON line: everything else (plugin banners, the NCCL version line that
    # the C library writes to fd 1) goes to stderr -- at the file-descriptor level.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(json_fd, 'w')
    sys.stdout = sys.stderr
    opt = AugOptions().parse(args={'p_thres': 0.0, 'w_lpips': 0.0, 'w_disc': args.w_disc, 'init_w': 'inv', 'n_imgs': 0}, argv=argv)
    aug = create_augment(opt)
    core = aug.latent_aug.module
    eng = core.engines[0]
    names = list(aug.stats_dataset_w.index.keys())

    def batch_data(i):
        fn = [names[(i * B + j) % len(names)] for j in range(B)]
        img = torch.zeros([B, 1, res, res])
        return {'A': img, 'B': img, 'A_paths': fn, 'B_paths': fn}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- kernel-resident metric: inputs already in HBM, no host copies in the timed region
    w_dev = [aug.sample_from_inversion(batch_data(i)['A_paths']).to(dev) for i in range(args.warmup + args.steps)]
    for i in range(args.warmup):
        core.forward(w_dev[i])
    clocks = ClockSampler(local_rank)
    barrier()
    mark = clocks.mark()
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        core.forward(w_dev[args.warmup + i])
    ev1.record()
    barrier()
    launches = eng.launch_count - l0
    ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    value = world * B / (ms * 1e-3)

    # ---- end to end through the plugin API: host dict in, host dict out
    if args.profile:
        clocks.stop(mark)
        if rank == 0:
            print(json.dumps({'metric': 'augmented images/sec', 'value': value, 'unit': 'img/s', 'ms_per_step': ms,
                              'gpu_launches': int(launches), 'note': 'profile mode (no e2e / roofline / cpu_baseline legs)'}),
                  file=real_stdout, flush=True)
        return
    time.sleep(3.0)      # both legs start from a comparable power / thermal state (the kernels are power-capped)
    for i in range(3):
        aug.set_input(batch_data(i)); aug.forward(); aug.get_output()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        aug.set_input(batch_data(i))
        aug.forward()
        out = aug.get_output()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e2e_s = max_over_ranks(t1 - t0) / args.steps
    clk = clocks.stop(mark)
    e2e = {'value': world * B / e2e_s, 'unit': 'img/s', 'h2d_bytes_per_step': B * eng.w_dim * 4,
           'd2h_bytes_per_step': B * C * res * res * 4}
    assert out['A'].shape == (B, 1, res, res) and bool(torch.isfinite(out['A']).all())

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (tap-GEMM): every launch of one Adam step timed alone
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except OSError:
            pass
        peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback 1.4 PFLOP/s sustained'
        rows, rgb_macs = layer_table(res, C, c.get('channel_base', 32768), c.get('channel_max', 512))
        t = eng.debug_time_gemms(reps=10)
        tot_ms = sum(t['forward']) + sum(t['dgrad'])
        alg = 2.0 * 2.0 * B * sum(r['alg_macs'] for r in rows)          # fwd + dgrad launches of one step
        exe = 2.0 * 2.0 * B * sum(r['exe_macs'] for r in rows) * (3 if args.precision == 'fp32_parity' else 1)
        ach = alg / (tot_ms * 1e-3) / 1e12
        traffic = None          # DRAM bytes (read + write) of the same 26 launches, from the committed ncu capture of this workload
        if args.config == 'c2' and args.precision == 'bf16':
            try:
                traffic = json.load(open(os.path.join(ROOT, 'profiles', 'r1c_tapgemm_traffic.json')))['dram_bytes_read_plus_write']
            except (OSError, KeyError, ValueError):
                pass
        roof = {'bound': 'tensor', 'kernel': 'tapgemm_kernel (26 launches of one Adam step: 13 forward + 13 data-gradient)',
                'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf, 'traffic': traffic,
                'peak_source': peak_src, 'executed_tflops': exe / (tot_ms * 1e-3) / 1e12,
                'launch_ms_sum': tot_ms, 'seed_ms': t['seed'], 'fir_pass_ms_sum': sum(t['fir_forward']) + sum(t['fir_backward']),
                'whole_path_frac': (value / world) * (2 * steps + 1) * f_syn(res, C, channel_base=c.get('channel_base', 32768),
                                                                            channel_max=c.get('channel_max', 512)) / (peak_tf * 1e12)}
        if args.layers_out:
            tab = [dict(r, fwd_ms=t['forward'][i], dgrad_ms=t['dgrad'][i], fir_fwd_ms=t['fir_forward'][i], fir_bwd_ms=t['fir_backward'][i],
                        fwd_alg_tflops=2.0 * B * r['alg_macs'] / (t['forward'][i] * 1e-3) / 1e12,
                        dgrad_alg_tflops=2.0 * B * r['alg_macs'] / (t['dgrad'][i] * 1e-3) / 1e12) for i, r in enumerate(rows)]
            json.dump({'config': args.config, 'precision': args.precision, 'batch': B, 'layers': tab, 'seed_ms': t['seed']},
                      open(args.layers_out, 'w'), indent=1)
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_baseline(args.config)
        line = {'metric': 'augmented images/sec', 'value': value, 'unit': 'img/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'bf16 operands, f32 accumulate' if args.precision == 'bf16' else 'split-bf16 (hi+lo) operands, f32 accumulate',
                'data': 'synthetic',
                'config': {'workload': workload_name(args.config, c) + (f', w_disc={args.w_disc:g} (StyleGAN2 discriminator term)' if args.w_disc > 0 else ''),
                           'precision': args.precision,
                           'l2': 'working set >> L2: ~2 GB of activations written and re-read per Adam step',
                           'parallelism': f'batch-sharded x{world}, no data-path collective'},
                'clocks': clk, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roof, 'cpu_baseline': cb}
    if rank == 0:
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
