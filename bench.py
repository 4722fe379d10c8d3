#!/usr/bin/env python
"""bench.py — weak-form residual + gradient throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one reference training step `sess.run([optMinimize, loss], feed)`
(VarNetUtility.py:1044): residual loss + all weight gradients (+ NCCL all-reduce at N>1) + TF-Adam
on one batch of synthetic input.  The workload at N=1 is BASELINE.json configs[3], the config the
metric is quoted on: 10^6 space-time test functions x 4^3 Gauss points (P = 6.4e7 quadrature
points), 4x64 tanh MLP, built with the reference API's discretisation (varnet_b200/workloads.py).
At N>1 the same 10^6 test functions are sharded contiguously over the ranks like the reference's
towers (VarNetUtility.py:830-838) => "strong" scaling; the only collective is the all-reduce of
the [grad | 4 loss scalars] buffer (51 KB).

Timing: W>=3 warm-up steps, then exactly K steps between barrier+synchronize, CUDA events on the
launching stream, max over ranks.  Inputs (1.5 GB/GPU at N=1) are far larger than L2 (126 MB), so no
explicit flush is needed.  `value` = resident-table throughput; `e2e` = the same step with the host feed re-uploaded
inside the timed region every step and the loss read back, at the reference's boundary: float64 NumPy arrays in pageable
memory (VarNetUtility.py:840-854) through vn_loss_grad_fed_f64 (host threads stage sub-chunks into pinned bounce buffers;
copies and the float64->float32 pack overlap the step's kernels); `e2e_f32_pinned` = the same with a pre-pinned float32 feed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nx, ny, ntime, layerWidth, activation)
    "synthetic_2dt_1e6x64_mlp4x64_tanh": (100, 100, 100, (64, 64, 64, 64), "tanh"),
    "synthetic_2dt_1e6x64_mlp4x16_tanh": (100, 100, 100, (16, 16, 16, 16), "tanh"),
    "synthetic_2dt_small_mlp4x64_tanh": (40, 40, 25, (64, 64, 64, 64), "tanh"),
    # width sweep (BASELINE.json configs[4]): widths above 64 run on the tensor-core class (tcgen05 3xTF32)
    "synthetic_2dt_1e6x64_mlp4x128_tanh": (100, 100, 100, (128, 128, 128, 128), "tanh"),
    "synthetic_2dt_1e6x64_mlp4x256_tanh": (100, 100, 100, (256, 256, 256, 256), "tanh"),
}
DEFAULT = "synthetic_2dt_1e6x64_mlp4x64_tanh"
METRIC = "weak-form residual+grad quad-pts/sec"
UNIT = "quad-pts/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT, choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--store-reference", action="store_true", help="N=1: write loss / ||grad|| of the seeded step to profiles/n1_reference.json")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-operator-configs", action="store_true", help="skip the steps/sec legs of configs 1-3")
    ap.add_argument("--no-width-sweep", action="store_true", help="skip the short width-16/128/256 legs (BASELINE config 5)")
    return ap.parse_args()


def measured_traffic(workload):
    """DRAM bytes per launch of the dominant kernel on this workload, from the committed
    `ncu` capture of this build (profiles/r2_traffic.json; regenerate with scripts/make_traffic_json.py)."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            t = json.load(f)
        if t.get("workload") == workload:
            return t.get("dram_bytes_per_launch")
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p.get("hbm_gbs", 6650.0), bf16_tflops=p.get("bf16_tflops_sustained", p.get("bf16_tflops", 2250.0)),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=2250.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []
        self.nvml, self.samples, self.stop_flag = None, [], False
        self.period = 0.02          # NVML queries take driver locks: keep the polling light (and on rank 0 only)

    def _poll_nvml(self):
        nv, h = self.nvml
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((float(sm), pw, int(rs)))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        # NVML polled from a thread (short timed regions at N=8 are over before nvidia-smi prints its first line)
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.nvml = (nv, h)
            self.th = threading.Thread(target=self._poll_nvml, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.th.join(timeout=2)
            bits = dict(sw_power_cap=0x4, hw_slowdown=0x8, sw_thermal_slowdown=0x20, hw_thermal_slowdown=0x40)
            reasons = sorted(k for k, b in bits.items() if any(r & b for _, _, r in self.samples))
            sm = [x[0] for x in self.samples]
            return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=self.smax, reasons=reasons,
                        power_w_max=max(x[1] for x in self.samples) if self.samples else None, samples=len(sm), source="nvml")
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2]); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=smax, reasons=sorted(reasons),
                    power_w_max=max(pw) if pw else None, samples=len(sm))


def bind_to_gpu_numa_node(gpu_index):
    """Pin this process to the CPUs NVML reports as local to its GPU (one process per GPU under torchrun): the pinned host
    feed of the e2e legs is then allocated on the NUMA node whose PCIe root the GPU hangs off, instead of wherever the
    launcher happened to start the process.  Best effort; returns the CPU count it bound to, or None."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def cpu_feed(nx, ny, ntime, sample_tf):
    """Bounded sample of the same workload for the CPU legs: the first `sample_tf` test functions."""
    from varnet_b200 import workloads
    feed, meta = workloads.shard_feed(nx, ny, ntime, 0, sample_tf, dtype=np.float32)
    return feed, meta


def run_cpu(args, nx, ny, ntime, lw, act, steps, warmup, budget_s=None):
    """The reference's CPU path for this step: torch CPU FP32 double back-prop mirroring the TF graph
    (oracle/torch_oracle.py; TF 1.10 is not installable), all host threads."""
    import torch
    from oracle import torch_oracle, graph_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_tf = (1024 if max(lw) > 64 else 4096) if len(lw) and max(lw) >= 64 else 16384   # 6.6e4 / 2.6e5 / 1.0e6 points per CPU step
    feed, meta = cpu_feed(nx, ny, ntime, sample_tf)
    theta = graph_oracle.glorot_init(meta["inpDim"], list(lw), seed=0)
    st = torch_oracle.CpuStepper(theta, feed, meta["dim"], meta["inpDim"], list(lw), act, True, meta["lossOpt"],
                                 threads=cores)
    P = sample_tf * meta["integNum"]
    for _ in range(warmup):
        st.step()
    times = []
    t_all = time.perf_counter()
    for k in range(steps):
        t0 = time.perf_counter()
        st.step()
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_all > budget_s and k >= 1:
            break
    dt = float(np.mean(times))
    # the float64->float32 feed cast the reference pays on every sess.run (VarNetUtility.py:1044)
    f64 = [np.asarray(feed[k], dtype=np.float64) for k in ("Input", "gcoef", "N", "dNt", "source")]
    t0 = time.perf_counter()
    for a in f64:
        a.astype(np.float32)
    cast = time.perf_counter() - t0
    return dict(value=P / dt, ms_per_step=dt * 1e3, cores=cores, steps=len(times), P=P, sample_tf=sample_tf,
                feed_cast_ms=cast * 1e3,
                sample="first %d of %d test functions (%d quad points) of the same workload, torch-CPU FP32 "
                       "double back-prop + TF-Adam, %d threads" % (sample_tf, nx * ny * ntime, P, cores))


def operator_cpu_baseline(feed, kw, seconds):
    """The reference's CPU path (torch FP32 double back-prop + TF-Adam, oracle/torch_oracle.py) on one feed dict of an
    operator config, with all host threads and with one thread."""
    import torch
    from oracle import torch_oracle, graph_oracle
    nb, q = [int(v) for v in feed["intShape"]]
    cap = 16384                                   # bounded sample for the 15.36 M-point config
    if nb > cap:
        feed = dict(feed)
        for k in ("Input", "gcoef", "source", "N", "dNt"):
            v = feed.get(k)
            if isinstance(v, np.ndarray) and v.dtype != object and v.shape[0] == nb * q:
                feed[k] = v[:cap * q]
        feed["intShape"] = [cap, q]
        nb = cap
    theta = graph_oracle.glorot_init(kw["inpDim"], kw["layerWidth"], seed=0)
    out = {}
    for label, threads in (("all_threads", os.cpu_count() or 1), ("one_thread", 1)):
        st = torch_oracle.CpuStepper(theta, feed, kw["dim"], kw["inpDim"], kw["layerWidth"], kw["activation"], kw["timeDependent"],
                                     kw["lossOpt"], threads=threads)
        st.step()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds or n < 2:
            st.step(); n += 1
        dt = (time.perf_counter() - t0) / n
        out[label] = dict(cores=threads, ms_per_step=dt * 1e3, steps_per_sec=1.0 / dt, quad_pts_per_sec=nb * q / dt)
    torch.set_num_threads(os.cpu_count() or 1)
    out["kind"] = "port"
    out["sample"] = "%d test functions x %d points of the config's first feed dict" % (nb, q)
    return out


def operator_config_steps(world=1, rank=0, budget_s=4.0, cpu_seconds=0.0):
    """train steps/sec of BASELINE.json configs 1-3 (the reference's operator scripts) on one GPU: each is
    built with the reference API of the host mirror and driven through ManageTrainData.optimIter ->
    TFNN.sess.run([optMinimize, loss]) exactly like VarNet.train's hot loop (VarNet.py:1350-1352)."""
    import varnet_b200
    from varnet_b200 import ManageTrainData
    pi = np.pi
    out = {}
    procs = 'GPU:0' if world == 1 else ['GPU:%d' % i for i in range(world)]

    def cfg1():
        pde = varnet_b200.ADPDE(varnet_b200.Domain1D(), diff=0.1 / pi, vel=1.0, timeDependent=True, tInterval=[0, 2.0],
                                IC=lambda x: -np.sin(pi * x))
        return varnet_b200.VarNet(pde, layerWidth=[20], discNum=20, bDiscNum=None, tDiscNum=300, processors=procs, seed=0), None, False

    def cfg2():
        v = np.array([[0.0, -0.5], [0.0, -0.2], [0.0, 0.2], [0.0, 0.5], [2.0, 0.5], [2.0, -0.5]])
        pde = varnet_b200.ADPDE(varnet_b200.PolygonDomain2D(v), diff=1.e-3, vel=[1., 0.], tInterval=[0, 1.5],
                                BCs=[[], [0.0, 1.0, 1.0], [], [], [], []], IC=0.0)
        return varnet_b200.VarNet(pde, layerWidth=[10, 20], discNum=[80, 40], bDiscNum=40, tDiscNum=75, processors=procs, seed=0), None, False

    def diffFun(x, t=0, D=0.1 / pi):
        return D * np.ones([np.shape(x)[0], 1])

    def cfg3():
        mor = varnet_b200.MOR(diffFun, ['D'], [[0.003, 0.033]])
        pde = varnet_b200.ADPDE(varnet_b200.Domain1D(), diff=diffFun, vel=1.0, timeDependent=True, tInterval=[0, 2.0],
                                IC=lambda x: -np.sin(pi * x), MORvar=mor)
        disc = lambda n=6: np.array([0.003 * (11 ** (k / (n - 1))) for k in range(n)])[np.newaxis].T
        return varnet_b200.VarNet(pde, layerWidth=[10, 20, 30], discNum=150, bDiscNum=75, tDiscNum=800, MORdiscScheme=disc,
                                  processors=procs, seed=0), 20, True

    for name, build in (("Operator_1Dt", cfg1), ("Operator_2Dt", cfg2), ("Operator_1DtMOR", cfg3)):
        t_build = time.perf_counter()
        vn, batchNum, saveMOR = build()
        fd = vn.fixData
        fd.setFEdata()
        Input, _, biInput, _ = vn.trainingPoints()
        disc = None if vn.PDE.MORvar is None else vn.PDE.MORvar.discretizeArg(vn.MORdiscScheme)
        tData = ManageTrainData(Input, biInput, batchNum, None, saveMOR, fd.MORbatchNum)
        tData = vn.trainData(0, disc, tData)
        tData.updateDictFields('trainW', np.array([10., 10., 1.]))
        t_build = time.perf_counter() - t_build
        tf = vn.tfData

        chunked = fd.MORbatchNum == 1 and tData.batchNum == 1     # what VarNet.train(stepsPerCall=64) does for such configs

        def epochs(n):
            """n epochs exactly as VarNet.train takes them; returns the number of optimizer steps."""
            nonlocal tData
            if chunked:
                done = 0
                while done < n:
                    k = min(64, n - done)
                    tData.optimIterMany(tf, k)
                    done += k
                return n
            steps = 0
            for _ in range(n):
                total = 0
                for b in range(fd.MORbatchNum):
                    tData = vn.trainData(b, disc, tData)
                    total += tData.optimIter(tf)        # may be a backend.Deferred: the steps are enqueued, not waited for
                    steps += tData.batchNum
                float(total)                            # the epoch's loss, as VarNet.train reads it: every step has completed
            return steps
        epochs(1)                                      # warm-up (also caches the MOR batches on the host)
        # every rank runs the same number of epochs (the steps contain a collective): fixed count from a short calibration
        n_cal = 64 if chunked else 1
        t0 = time.perf_counter()
        epochs(n_cal)
        t_cal = (time.perf_counter() - t0) / n_cal
        n_epochs = max(1, int(budget_s / max(t_cal, 1e-7)))
        if world > 1:
            import torch
            import torch.distributed as dist
            box = torch.tensor([n_epochs], dtype=torch.int64, device="cuda")
            dist.broadcast(box, src=0)
            n_epochs = int(box[0])
            dist.barrier()
        t0 = time.perf_counter()
        steps = epochs(n_epochs)
        dt = time.perf_counter() - t0
        tw_loc = next(t for t in tf.compTowers if t.local)
        P = int(np.prod(tData.optimFeedicts[0][tw_loc.intShape])) * world
        out[name] = dict(steps_per_sec=steps / dt, quad_points_per_step=P, quad_pts_per_sec=steps * P / dt,
                         steps_per_epoch=int(fd.MORbatchNum * tData.batchNum), table_build_s=t_build, n_gpus=world,
                         host_round_trips="one per 64 steps (vn_train_steps)" if chunked else
                         ("one per %d mini-batch steps (vn_train_batches_begin/_end; the host prepares the next MOR batch meanwhile)" % tData.batchNum
                          if tData.batchNum > 1 else "one per step"))
        if cpu_seconds > 0 and rank == 0:
            f0 = {k.name: (np.array(v) if type(v).__name__ == "TableView" else v) for k, v in tData.optimFeedicts[0].items()
                  if getattr(k, "tower", None) == tw_loc.index}
            kw = dict(dim=tf.dim, inpDim=tf.inpDim, layerWidth=list(tf.layerWidth), activation="sigmoid", timeDependent=tf.timeDependent,
                      lossOpt=tf.lossOpt)
            try:
                out[name]["cpu_baseline"] = operator_cpu_baseline(f0, kw, cpu_seconds)
            except Exception as ex:
                out[name]["cpu_baseline"] = dict(error=repr(ex))
        tf.sess.close()
    return out


def width_sweep(feed, meta, act, P_local, widths=(16, 64, 128, 256), steps=2):
    """BASELINE.json config 5 on the same mesh and feed: depth-4 tanh MLPs of other widths, a few steps each
    (resident tables, CUDA events).  Widths above 64 run on the tensor-core class (tcgen05 3xTF32); the width-64 leg
    is the FP32-FMA tile kernel (VARNET_B200_CLASS=fma) for the FMA-vs-tensor-core crossover at the headline width —
    the main line of this run is the same network on the tensor-core tile kernel."""
    import torch
    from varnet_b200 import workloads
    from varnet_b200.backend import TFNN
    out = {}
    for wdt in widths:
        lw = [wdt] * 4
        prev = os.environ.get("VARNET_B200_CLASS")
        if wdt == 64:
            os.environ["VARNET_B200_CLASS"] = "fma"
        tf = TFNN(meta["dim"], meta["inpDim"], lw, "MLP", act, True, None, ["GPU:0"], None, meta["lossOpt"], "adam", 1e-3, seed=0)
        tw = tf.compTowers[0]
        fd = {getattr(tw, k): feed[k] for k in ("Input", "gcoef", "source", "N", "dNt", "biInput", "biLabel", "bDof",
                                                 "intShape", "integW", "biDimVal", "detJvec", "detJ", "w")}
        tf.sess.run([tf.optMinimize, tf.loss], feed_dict=fd)                 # upload + warm-up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            loss = tf.sess.run([tf.optMinimize, tf.loss], feed_dict=fd)[1]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        flop_pt = workloads.algorithmic_flops_per_point(meta["inpDim"], meta["dim"], lw)
        if wdt == 64:
            if prev is None:
                os.environ.pop("VARNET_B200_CLASS", None)
            else:
                os.environ["VARNET_B200_CLASS"] = prev
        out["mlp4x%d%s" % (wdt, "_fma" if wdt == 64 else "")] = dict(quad_pts_per_sec=P_local / (ms * 1e-3), ms_per_step=ms, algorithmic_tflops=flop_pt * P_local / (ms * 1e-3) / 1e12,
                                    loss=float(loss), kernel_family=tw.engine.kernel_info().split()[0])
        tf.sess.close()
    return out


def main():
    args = parse()
    nx, ny, ntime, lw, act = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = dict(workload=args.workload, test_functions=nx * ny * ntime, integNum=64, quad_points=nx * ny * ntime * 64,
                  mlp="%dx%d %s" % (len(lw), lw[0], act), inpDim=3, dim=2, sharding="contiguous test-function ranges per rank",
                  l2_policy="per-GPU tables larger than L2, no flush")

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = run_cpu(args, nx, ny, ntime, lw, act, args.steps, args.warmup)
        line = dict(metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=r["steps"], warmup=args.warmup,
                    ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype="f32", data="synthetic", impl="reference", config=config,
                    cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"],
                                      feed_cast_ms_per_step=r["feed_cast_ms"]),
                    e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0)
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the varnet_b200 engine has no CPU path")
    torch.cuda.set_device(local_rank)
    # N > 1: pinned staging buffers are first-touched on the GPU's own NUMA node (N = 1 keeps all cores for the CPU-baseline legs)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from varnet_b200 import workloads
    from varnet_b200.backend import TFNN
    from varnet_b200._capi import fp32_peak_tflops

    # ---- this rank's tower: contiguous range of test functions
    nt = nx * ny * ntime
    n0, n1 = workloads.tower_range(nt, world, rank)
    t_build = time.perf_counter()
    feed, meta = workloads.shard_feed(nx, ny, ntime, n0, n1, dtype=np.float32)
    # BC/IC rows are replicated on every tower and down-weighted by 1/puNum (VarNetUtility.py:900-901)
    feed["w"] = np.array([1.0 / world, 1.0 / world, 1.0])
    pinned = {}
    for k in ("Input", "gcoef", "dNt", "biInput", "biLabel"):
        t = torch.from_numpy(np.ascontiguousarray(feed[k])).pin_memory()
        pinned[k] = t
        feed[k] = t.numpy()
    t_build = time.perf_counter() - t_build
    procs = ["GPU:%d" % i for i in range(world)]
    tf = TFNN(meta["dim"], meta["inpDim"], list(lw), "MLP", act, True, None, procs, None, meta["lossOpt"], "adam", 1e-3, seed=0)
    tw = tf.compTowers[rank]
    eng = tw.engine
    fd = {getattr(tw, k): feed[k] for k in ("Input", "gcoef", "source", "N", "dNt", "biInput", "biLabel", "bDof",
                                             "intShape", "integW", "biDimVal", "detJvec", "detJ", "w")}
    P_local = (n1 - n0) * meta["integNum"]
    P_total = nt * meta["integNum"]
    h2d = sum(feed[k].nbytes for k in ("Input", "gcoef", "dNt", "biInput", "biLabel"))

    def step():
        return tf.sess.run([tf.optMinimize, tf.loss], feed_dict=fd)[1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity of the step itself, before any update: loss and ||grad|| of the seeded weights on this table, summed over
    # the towers.  At N = 1 they are compared with the stored reference (profiles/n1_reference.json, written by an N = 1 run
    # with --store-reference); at N > 1 a deviation above 1e-6 relative fails the run (exit code 3): the all-reduced
    # N-tower step must be the 1-tower step.
    g0, loss0 = tf.sess.run([tf.grad, tf.loss], feed_dict=fd)
    gnorm0 = float(np.linalg.norm(g0.astype(np.float64)))
    parity = dict(loss=float(loss0), grad_norm=gnorm0)
    ref_path = os.path.join(ROOT, "profiles", "n1_reference.json")
    stored = {}
    if os.path.exists(ref_path):
        with open(ref_path) as f:
            stored = json.load(f)
    parity_fail = False
    if args.workload in stored:
        r0 = stored[args.workload]
        parity["stored_n1"] = r0
        parity["rel_diff_loss"] = abs(float(loss0) - r0["loss"]) / abs(r0["loss"])
        parity["rel_diff_grad_norm"] = abs(gnorm0 - r0["grad_norm"]) / abs(r0["grad_norm"])
        parity_fail = world > 1 and max(parity["rel_diff_loss"], parity["rel_diff_grad_norm"]) > 1e-6
    if args.store_reference and world == 1 and rank == 0:
        stored[args.workload] = dict(loss=float(loss0), grad_norm=gnorm0, kernel=eng.kernel_info().split()[0])
        with open(ref_path, "w") as f:
            json.dump(stored, f, indent=1, sort_keys=True)

    for _ in range(max(args.warmup, 3)):
        loss = step()
    # ---- timed region: the PRODUCT step (CUDA graph, optimizer fused into the reduction at N = 1, all-reduce inside the
    # graph at N > 1, boundary/initial rows on the auxiliary stream).  Every step also gets its own event pair so that a
    # straggling rank or step can be named.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()            # before the barrier: starting the NVML thread takes milliseconds that the other ranks would
                                   # otherwise spend waiting for rank 0 inside the first timed all-reduce
        time.sleep(0.05)
    barrier()
    l0 = eng.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for k in range(args.steps):
        loss = step()
        evs[k + 1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = evs[0].elapsed_time(evs[-1])
    step_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    launches = eng.launch_count() - l0

    # ---- profiled leg (separate, not the timed region): per-kernel CUDA events inside the engine; plain launches, the
    # optimizer as separate kernels, boundary/initial rows in sequence
    eng.profile_enable(True)
    eng.profile_read()
    barrier()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(min(args.steps, 3)):
        loss = step()
    pe1.record()
    barrier()
    ms_prof = pe0.elapsed_time(pe1) / min(args.steps, 3)
    prof = eng.profile_read()
    eng.profile_enable(False)

    # ---- the collective alone: K all-reduces of the [grad | 4 scalars] buffer back to back (N > 1, native communicator)
    ms_comm = None
    if world > 1 and tf.native_comm:
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(20):
            eng.allreduce_grad()
        eng.synchronize()
        c1.record()
        torch.cuda.synchronize()
        ms_comm = c0.elapsed_time(c1) / 20

    kernel_ms_rank = prof["var_adj"][0] / max(prof["var_adj"][1], 1)
    per_rank = dict(rank=rank, total_ms=ms, step_ms=step_ms, kernel_ms=kernel_ms_rank, allreduce_ms=ms_comm)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, per_rank)
        t = torch.tensor([ms, float(launches)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, launches = float(tmax[0]), int(tsum[1])
    else:
        gathered = [per_rank]
    ms_step = ms / args.steps

    # ---- e2e legs: the host feed is re-uploaded inside the timed region every step and the loss is read back.
    #   f32_pinned: float32 arrays in pinned host memory (the friendliest caller);
    #   f64_pageable: what the reference's callers hand over (VarNetUtility.py:840-854): float64 NumPy arrays in pageable
    #   memory, cast to float32 on the device (vn_loss_grad_fed_f64).
    def e2e_leg(feed_dict, nsteps):
        tf.feed_cache = False
        step_fd = lambda: tf.sess.run([tf.optMinimize, tf.loss], feed_dict=feed_dict)[1]
        for _ in range(3):                        # warm-up; arrays fed for the second time are page-locked in place (_capi.HostPins)
            step_fd()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(nsteps):
            lv = step_fd()
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        m = max(e0.elapsed_time(e1), wall)            # host-side staging (pageable copies) is part of the step
        if world > 1:
            te = torch.tensor([m], dtype=torch.float64, device="cuda")
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            m = float(te[0])
        tf.feed_cache = True
        return m / nsteps, float(lv)

    ms_e2e, _ = e2e_leg(fd, args.e2e_steps)
    fd64 = dict(fd)
    for k in ("Input", "gcoef", "dNt", "biInput", "biLabel"):
        fd64[getattr(tw, k)] = np.asarray(feed[k], dtype=np.float64)            # pageable float64 copies
    h2d64 = sum(fd64[getattr(tw, k)].nbytes for k in ("Input", "gcoef", "dNt", "biInput", "biLabel"))
    ms_e2e64, _ = e2e_leg(fd64, args.e2e_steps)
    pins = tf.sess._host_pins() if hasattr(tf.sess, "_host_pins") else None
    host_registered = (dict(arrays=int(pins.registered), bytes=int(pins.bytes), cap_bytes=int(pins.cap), error=pins.error, touches=int(pins.clock))
                       if pins is not None else None)
    del fd64

    # the same table built on the device from the mesh centres + periodic FE tables (vn_generate_table_f64): compare
    # with table_build_s (host NumPy build + pinning); spare slot, freed again
    t_gen, gen_step = None, None
    try:
        eng.select_table(1)
        t0g = time.perf_counter()
        workloads.generate_on_device(eng, nx, ny, ntime, n0, n1)
        eng.synchronize()
        t_gen = time.perf_counter() - t0g
        # the same step on the generated table: for the 33-64-wide class nothing of size nT exists on the device, the tile kernel
        # rebuilds each row from the centre of its test function (kernel_info: table=in-kernel generation)
        if world == 1:
            for _ in range(2):
                eng.train_step(1e-3)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(3):
                eng.train_step(1e-3, fetch_loss=False)
            eng.synchronize()
            g1.record()
            torch.cuda.synchronize()
            ms_g = g0.elapsed_time(g1) / 3
            gen_step = dict(ms_per_step=ms_g, quad_pts_per_sec=P_local / (ms_g * 1e-3), in_kernel="in-kernel" in eng.kernel_info(),
                            table_bytes_per_point=(0.4 if "in-kernel" in eng.kernel_info() else 24))
        eng.free_table(1)
        eng.select_table(0)
    except Exception:
        t_gen = None
    if rank == 0:
        flop_pt = workloads.algorithmic_flops_per_point(meta["inpDim"], meta["dim"], lw)
        adj_ms, adj_n = prof["var_adj"]
        fwd_ms, fwd_n = prof["var_fwd"]
        kernel_ms = adj_ms / max(adj_n, 1)
        # the adjoint kernel recomputes the forward sweep and does both adjoint GEMMs per layer: its
        # algorithmic work is the whole residual+gradient count F_alg minus nothing (the separate
        # forward pass that produces R_i is extra, non-algorithmic work and is NOT credited)
        peak_tf = fp32_peak_tflops(local_rank)
        achieved = flop_pt * P_local / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0
        pk = peaks()
        bytes_pt = 4 * (meta["inpDim"] + meta["dim"] + 1)
        roofline = dict(bound="fp32", achieved=achieved, peak=peak_tf, unit="TFLOP/s", frac=achieved / peak_tf if peak_tf else None,
                        traffic=measured_traffic(args.workload) if world == 1 else None, kernel="vn_adj_kernel<MODE_VAR_FUSED>", kernel_ms=kernel_ms,
                        kernel_share_of_step=kernel_ms / ms_prof if ms_prof else None,     # both from the profiled leg (plain launches, per-kernel events)
                        flop_per_point=flop_pt, points_per_launch=P_local,
                        peak_source="FFMA microbenchmark (vn_fp32_peak_tflops) measured in this run",
                        hbm=dict(bound="hbm", achieved=bytes_pt * P_local / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0,
                                 peak=pk["hbm_gbs"], unit="GB/s", bytes_per_point=bytes_pt, peak_source=pk["source"]),
                        other_kernels_ms=dict(var_fwd=fwd_ms / max(fwd_n, 1), segreduce=prof["segreduce"][0] / max(prof["segreduce"][1], 1),
                                              bic=prof["bic"][0] / max(prof["bic"][1], 1), finalize=prof["finalize"][0] / max(prof["finalize"][1], 1),
                                              optimizer=prof["optimizer"][0] / max(prof["optimizer"][1], 1)))
        roofline["hbm"]["frac"] = roofline["hbm"]["achieved"] / pk["hbm_gbs"]
        if "tcgen05" in eng.kernel_info():
            # tensor-core class: the layer GEMMs run as 3xTF32 (three tcgen05.mma kind::tf32 per algorithmic product);
            # the bound is the TF32 tensor peak = half the measured dense bf16 figure (sustained: kernel timed inside a long step)
            tf32_peak = 0.5 * pk["bf16_tflops"]
            tile64 = "tile64" in eng.kernel_info()
            roofline.update(bound="tensor", peak=tf32_peak, frac=achieved / tf32_peak,
                            kernel="tc64_var_kernel (vn_tc64.cu: resident 128-point tiles, A-from-TMEM 3xTF32)" if tile64 else "tc_gemm_kernel / tc_gw_kernel pipeline (vn_tc.cu)",
                            peak_source="0.5 x dense bf16 (%s): TF32 runs at half the bf16 rate" % pk["source"],
                            executed_mma_tflops=3.0 * achieved, executed_frac=3.0 * achieved / tf32_peak,
                            fp32_fma_peak=peak_tf, frac_of_fp32_fma_peak=achieved / peak_tf if peak_tf else None)
        all_steps = np.array([g["step_ms"] for g in gathered])                      # [rank][step]
        slow = int(np.argmax([g["total_ms"] for g in gathered]))
        ranks = dict(step_ms_min=float(all_steps.min()), step_ms_median=float(np.median(all_steps)), step_ms_max=float(all_steps.max()),
                     per_rank_mean_ms=[float(np.mean(g["step_ms"])) for g in gathered], slowest_rank=slow,
                     per_rank_kernel_ms=[float(g["kernel_ms"]) for g in gathered],
                     allreduce_ms=(None if gathered[0]["allreduce_ms"] is None else float(max(g["allreduce_ms"] for g in gathered))),
                     comm="native NCCL communicator inside the step graph (vn_comm_init)" if getattr(tf, "native_comm", False) else
                          ("torch.distributed all_reduce" if world > 1 else "none"))
        line = dict(metric=METRIC, value=P_total / (ms_step * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms_step, higher_is_better=True, scaling="strong",
                    vs_baseline=None, dtype="f32", data="synthetic", config=config,
                    train_steps_per_sec=1e3 / ms_step, loss=float(loss), parity=parity, table_build_s=t_build, device_table_generate_s=t_gen, generated_table_step=gen_step,
                    timed_path="product step: one CUDA graph per step (kernels, %s optimizer)" % ("all-reduce," if world > 1 else "fused"),
                    profiled_ms_per_step=ms_prof, ranks=ranks,
                    roofline=roofline,
                    # headline e2e = the reference boundary: float64 NumPy arrays in pageable memory, as VarNet's callers feed them
                    e2e=dict(value=P_total / (ms_e2e64 * 1e-3), unit=UNIT, h2d_bytes_per_step=int(h2d64), d2h_bytes_per_step=4,
                             ms_per_step=ms_e2e64, steps=args.e2e_steps,
                             feed="float64 NumPy arrays in pageable host memory, as the reference's callers feed them (VarNetUtility.py:840-854)",
                             warmup=3, host_registered=host_registered,
                             api="TFNN.sess.run([optMinimize, loss], feed_dict) with feed_cache=False -> vn_loss_grad_fed_f64: the first step stages "
                                 "sub-chunks through pinned bounce buffers with host threads; arrays fed a second time are page-locked in place "
                                 "(vn_host_register) and copied by DMA as float64, cast by the pack kernel; copies overlap the step's kernels"),
                    e2e_f32_pinned=dict(value=P_total / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=4,
                                        ms_per_step=ms_e2e, steps=args.e2e_steps, feed="float32 arrays in pinned host memory (the friendliest caller)",
                                        api="TFNN.sess.run([optMinimize, loss], feed_dict) with feed_cache=False -> vn_loss_grad_fed_f32"),
                    gpu_launches=int(launches), clocks=clocks, kernel_info=eng.kernel_info(), cpus_bound_to_gpu_numa_node=numa)
        if world == 1 and not args.no_cpu_baseline:
            r = run_cpu(args, nx, ny, ntime, lw, act, steps=50, warmup=1, budget_s=args.cpu_seconds)
            line["cpu_baseline"] = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"],
                                        ms_per_step=r["ms_per_step"], feed_cast_ms_per_step=r["feed_cast_ms"])
        if world == 1 and not args.no_width_sweep and args.workload == DEFAULT:
            try:
                line["width_sweep"] = width_sweep(feed, meta, act, P_local)
            except Exception as ex:
                line["width_sweep"] = dict(error=repr(ex))
    tf.sess.close()
    # ---- train steps/sec of the reference's own three operator configurations (BASELINE.json configs 1-3) at this N: every
    # rank takes part (its tower of each config), rank 0 reports.  They are launch-bound and do not scale: reported, not hidden.
    opcfg = None
    if not args.no_operator_configs:
        try:
            opcfg = operator_config_steps(world, rank, budget_s=4.0 if world == 1 else 2.0,
                                          cpu_seconds=(0.0 if (args.no_cpu_baseline or world > 1) else 2.0))
        except Exception as ex:                          # never lose the headline line to an auxiliary leg
            opcfg = dict(error=repr(ex))
    if rank == 0:
        if opcfg is not None:
            line["operator_configs"] = opcfg
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity_fail:
        sys.stderr.write("bench.py: N=%d step differs from the stored N=1 step: %s\n" % (world, json.dumps(parity)))
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
