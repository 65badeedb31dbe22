#!/usr/bin/env python
"""bench.py — sampled edges/s of the factored sampling-and-extraction hot path.

Workload (BASELINE.json configs[1]): GraphSAGE 2-layer fanout [25,10], batch 8000,
khop2 sampler, PreSC feature cache, on a papers100M-shaped synthetic power-law
graph (111M nodes, 1.6B edges, 128-d fp32 features), one B200 per rank.

A "step" is one mini-batch of the hot path: shuffle slice -> k-hop sample ->
ordered unique -> remap (per layer) -> cache lookup + feature gather + label
gather.  `value` = sampled edges/s with everything resident in HBM (device
timed); `e2e` = the same metric through the samgraph_* C-ABI host runtime with
the per-step host<->device traffic inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "fgnn-artifacts_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "sampled edges/s (sample + unique/remap + cache-aware extract per mini-batch)"
UNIT = "edges/s"
SEED_WEIGHTS = 0x46474E50
DTYPE = "u32"          # ids, offsets and hashes are uint32 arithmetic; fp32 feature rows are moved as bytes
FANOUTS = [25, 10]
BATCH = 8000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1510)     # ten papers100M epochs at batch 8000 (151 steps each)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FGNN_BENCH_WORKLOAD", "papers100M"))
    ap.add_argument("--cache-pct", type=float, default=float(os.environ.get("FGNN_BENCH_CACHE_PCT", "1.0")),
                    help="fraction of vertices whose features are cached in HBM (PreSC order); 1.0 = the whole\n                    57 GB table is HBM-resident on a 180 GB B200; the reference-like 25%% regime is always\n                    measured too and reported under extra.cache25")
    ap.add_argument("--empty-feat", type=int, default=int(os.environ.get("FGNN_BENCH_EMPTY_FEAT", "22")),
                    help="host feature table has 2^k rows, indices masked (SAMGRAPH_EMPTY_FEAT semantics)")
    ap.add_argument("--super", dest="super_batch", type=int, default=int(os.environ.get("FGNN_BENCH_SUPER", "4")),
                    help="mini-batches per super-batch: one fgnn_k_sample_batch_multi call samples them together "
                         "(two launches per layer for all of them); two super-batches alternate, one being sampled "
                         "while the other is extracted (device-resident leg; the engine does the same)")
    ap.add_argument("--sample-type", default=os.environ.get("FGNN_BENCH_SAMPLE_TYPE", "khop2"),
                    choices=["khop2", "khop0", "khop1", "weighted_khop", "weighted_khop_prefix",
                             "weighted_khop_hash_dedup", "random_walk"],
                    help="sampler of the device-resident leg (BASELINE configs #3-#5); anything but the default khop2 "
                         "skips the e2e and CPU-baseline legs, which are defined for the headline configuration only")
    ap.add_argument("--fanout", default=os.environ.get("FGNN_BENCH_FANOUT", "25,10"),
                    help="samgraph fanout list (sampled last to first), e.g. 5,10,15 for GCN; random_walk ignores it "
                         "(3 layers x top-5 of 4 walks of length 3, restart 0.5: train_pinsage.py:122-126)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-factored", action="store_true",
                    help="skip the factored (arch5: sampler GPUs + trainer GPUs) e2e leg and the measured epoch")
    ap.add_argument("--no-cache25", action="store_true", help="skip the extra 25 %% cache leg (profiling runs)")
    ap.add_argument("--only-partition", action="store_true",
                    help="diagnosis: N > 1, only the device-resident replicated + partitioned legs")
    ap.add_argument("--no-partition", action="store_true",
                    help="N > 1: skip the partitioned-cache legs (value is then the replicated-cache arm)")
    ap.add_argument("--replicate-pct", type=float, default=float(os.environ.get("FGNN_BENCH_REPLICATE_PCT", "0.75")),
                    help="N > 1, partitioned cache: fraction of the vertices (hottest PreSC ranks) that every GPU "
                         "keeps a copy of; the rest of the cache is striped over the GPUs and read by NVLink peer loads")
    return ap.parse_args()


# ---------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region.  NVML in a background thread (5 ms period:
    the default timed region lasts ~0.15 s, too short for `nvidia-smi -lms`, whose first sample arrives after its
    own start-up); nvidia-smi is the fallback when the NVML binding is unavailable."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.p = None
        self.f = None
        self.thread = None
        self.stop_flag = False
        self.recording = False       # samples are kept only between begin() and stop(): the timed region
        self.nv = self.h = None
        self.sm, self.mx, self.reasons = [], [], set()

    def _nvml_handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.idx)

    def _sample(self, nv, h, mx):
        try:
            sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            r = int(get_reasons(h))
        except Exception:
            return
        self.sm.append(sm)
        self.mx.append(float(mx))
        for bit, name in self.REASONS:
            if r & bit:
                self.reasons.add(name)

    def _loop(self, nv, h, mx):
        while not self.stop_flag:
            if self.recording:
                self._sample(nv, h, mx)
            time.sleep(0.005 if self.recording else 0.001)

    def start(self):
        """Bring the sampler up (NVML init + thread start take tens of ms: do it before the timed region);
        nothing is recorded until begin()."""
        try:
            import threading
            nv, h = self._nvml_handle()
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.nv, self.h, self.mx_clock = nv, h, mx
            self.thread = threading.Thread(target=self._loop, args=(nv, h, mx), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.QUERY,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def begin(self):
        self.recording = True

    def _one_shot_smi(self):
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.QUERY,
                                "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10)
            c = [x.strip() for x in r.stdout.strip().splitlines()[0].split(",")]
            self.sm.append(float(c[1]))
            self.mx.append(float(c[2]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)
        except Exception:
            pass

    def stop(self):
        if self.thread is not None:
            if not self.sm:      # region shorter than one period: one sample right at its end, still under load
                self._sample(self.nv, self.h, self.mx_clock)
            self.recording = False
            self.stop_flag = True
            self.thread.join(timeout=2)
            source = "nvml, 5 ms period, timed region only"
            if not self.sm:      # NVML gave nothing: one nvidia-smi query, taken right after the region
                self._one_shot_smi()
                source = "nvidia-smi, one query at the end of the timed region"
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None,
                    "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": source}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def step_of(k, rank, world, steps_per_epoch):
    """Mini-batch index (within an epoch) that rank `rank` processes at its k-th step: consecutive batches are
    dealt round-robin, so ranks never touch the same batch and no collective is needed on the data path."""
    return (k * world + rank) % steps_per_epoch


def aggregate(ms_total, edges, device):
    """Whole-job numbers: time = max over ranks, work = sum over ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms_total, edges
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e = torch.tensor([edges], dtype=torch.int64, device=device)
    dist.all_reduce(e, op=dist.ReduceOp.SUM)
    return float(t.item()), int(e.item())


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this same
    command (profiles/gather_traffic.json, written by tools/ncu_summary.py traffic); None when absent."""
    p = os.path.join(ROOT, "profiles", "gather_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


RW = dict(random_walk_length=3, random_walk_restart_prob=0.5, num_random_walk=4, num_neighbor=5, num_layer=3)


def fanouts_of(args):
    """Per-layer fanout of the device-resident leg; the default is FANOUTS (GraphSAGE [25,10])."""
    if getattr(args, "sample_type", "khop2") == "random_walk":
        return [RW["num_neighbor"]] * RW["num_layer"]
    return [int(x) for x in str(getattr(args, "fanout", "25,10")).split(",") if x]


def is_headline(args):
    return getattr(args, "sample_type", "khop2") == "khop2" and fanouts_of(args) == FANOUTS


def workload_name(args):
    """config.workload, identical for both arms (--impl ours / reference)."""
    if is_headline(args):
        return "GraphSAGE [25,10] batch 8000 khop2, %s-shaped synthetic power-law graph, PreSC cache %.0f%%" \
            % (args.workload, args.cache_pct * 100)
    return "fanout %s batch 8000 %s, %s-shaped synthetic power-law graph, PreSC cache %.0f%%" \
        % (fanouts_of(args), args.sample_type, args.workload, args.cache_pct * 100)


def shared_config(args, V, E, D, host_rows):
    """config: byte-identical in both arms (--impl ours / reference); arm-specific remarks go to `notes`."""
    return {"workload": workload_name(args), "batch": BATCH, "fanout": fanouts_of(args),
            "sample_type": args.sample_type, "num_node": V, "num_edge": E, "feat_dim": D,
            "synthetic_degree_exponent": 0.5, "cache_percentage": args.cache_pct,
            "e2e_cache_percentage": E2E_CACHE_PCT, "host_feat_rows": host_rows,
            "l2_policy": "inputs larger than L2 (GBs of topology and cache); no flush needed"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# ---------------------------------------------------------------------------
def build_workload(args, device):
    """Graph + labels + train set on `device`; host feature table (2^k rows, pinned)."""
    import torch
    from fgnn_b200.synth import SHAPES, SEED, make_graph_torch

    V, E, D, C, T = SHAPES[args.workload]
    t0 = time.time()
    indptr, indices = make_graph_torch(V, E, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(SEED + 1)
    label = torch.randint(0, C, (V,), generator=g, device=device, dtype=torch.int64)
    train = torch.randperm(V, generator=g, device=device)[:T].to(torch.int32)
    rows = min(V, 1 << args.empty_feat)
    gh = torch.Generator()
    gh.manual_seed(SEED + 1)
    host_feat = torch.rand((rows, D), generator=gh, dtype=torch.float32) * 2 - 1
    if torch.cuda.is_available():      # the reference arm also runs on a box without a GPU
        host_feat = host_feat.pin_memory()
    mask = rows - 1 if rows < V else 0xFFFFFFFFFFFFFFFF
    if rows < V:
        assert rows & (rows - 1) == 0
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return dict(V=V, E=E, D=D, C=C, T=T, indptr=indptr, indices=indices, label=label, train=train,
                host_feat=host_feat, feat_mask=mask, gen_s=time.time() - t0)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from fgnn_b200 import kernels as K
    from fgnn_b200.pipeline import HotPath

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout for the ONE json line (NCCL_DEBUG=VERSION boxes)
        dist.init_process_group("nccl", device_id=torch.device(dev))
    K.load()  # raises when the CUDA extension is missing: there is no CPU fallback

    wl = build_workload(args, dev)
    V, D = wl["V"], wl["D"]
    row_bytes = D * 4
    fanouts = fanouts_of(args)
    tables = {}
    if args.sample_type.startswith("weighted"):
        # kDefault weights of the reference's tools (integers 1..10, create_alias_table.cc:113), tables built on the GPU
        E = wl["E"]
        gw = torch.Generator(device=dev)
        gw.manual_seed(SEED_WEIGHTS)
        weights = torch.randint(1, 11, (E,), generator=gw, device=dev, dtype=torch.int32).to(torch.float32)
        if args.sample_type == "weighted_khop_prefix":
            tables["prefix_table"] = torch.empty(E, dtype=torch.float32, device=dev)
            K.build_prefix_table(wl["indptr"], V, weights, tables["prefix_table"])
        else:
            tables["prob_table"] = torch.empty(E, dtype=torch.float32, device=dev)
            tables["alias_table"] = torch.empty(E, dtype=torch.int32, device=dev)
            K.build_alias_table(wl["indptr"], wl["indices"], V, E, weights, tables["prob_table"], tables["alias_table"])
        torch.cuda.synchronize()
        del weights
        torch.cuda.empty_cache()
    hp = HotPath(wl["indptr"], wl["indices"], V, fanouts, BATCH, args.sample_type, seed=0x5EED0000 + rank, device=dev,
                 num_slots=2 * args.super_batch, rw=RW if args.sample_type == "random_walk" else None, **tables)
    steps_per_epoch = (wl["T"] + BATCH - 1) // BATCH
    # DistShuffler split (dist_shuffler.cc:60-83): rank r owns a contiguous range of the epoch's steps
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    perm = wl["train"][torch.randperm(wl["T"], generator=g, device=dev)].contiguous()

    def seeds_of(step):
        s = step_of(step, rank, world, steps_per_epoch)
        lo = s * BATCH
        hi = min(wl["T"], lo + BATCH)
        return perm[lo:hi], hi - lo

    # ---- PreSC: one pre-sampling epoch -> hotness ranking -> cache (cuda/pre_sampler.cc:57-110).
    # N > 1: every rank pre-samples its own share of the epoch's mini-batches, the visit counters are summed with
    # one NCCL all-reduce, rank 0 ranks the vertices and the ranking is broadcast once (the reference publishes
    # sampler 0's ranking through shared memory, dist_engine.cc:119-123).  No collective after this point.
    from fgnn_b200 import partition as P
    t0 = time.time()
    freq = torch.zeros(V, dtype=torch.int32, device=dev)
    for s in range(rank, steps_per_epoch, world):
        lo = s * BATCH
        hi = min(wl["T"], lo + BATCH)
        hp.sample(perm[lo:hi], hi - lo, 1_000_000 + s)
        hp.presample_count(freq)
    P.allreduce_freq(freq)
    rank_nodes = torch.empty(V, dtype=torch.int32, device=dev)
    if rank == 0:
        wsr = torch.empty(K.presc_rank_workspace_bytes(V), dtype=torch.uint8, device=dev)
        K.presc_rank(freq, V, rank_nodes, wsr)
        torch.cuda.synchronize()
        del wsr
    P.broadcast_ranking(rank_nodes, src=0)
    torch.cuda.synchronize()
    del freq
    presc_s = time.time() - t0
    hp.set_labels(wl["label"])

    G = args.super_batch
    S = len(hp.slots)                      # two slot groups of G mini-batches
    s_streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    x_stream = torch.cuda.Stream(device=dev, priority=-1)

    def measure(cache_pct, Ksteps, W, key0, profile=False, partition=False, replicate_pct=None):
        """Build the cache at `cache_pct` (replicated per GPU, or with partition=True striped over the ranks'
        GPUs and read through NVLink peer mappings), run W warm-up + Ksteps timed steps; device-timed.
        Like the engine's pump: G mini-batches are sampled together by ONE fgnn_k_sample_batch_multi call on the
        slot group's stream while the extraction stream gathers the features of the previous group's batches one
        after the other; a slot group is resampled only after its batches have been gathered."""
        t0 = time.time()
        hp.cache = None
        hp.feat_out = None
        torch.cuda.empty_cache()
        shards = None
        if partition:
            n_cached = int(V * cache_pct)
            n_repl = int(V * replicate_pct) if replicate_pct else 0
            n_repl = min(n_repl, n_cached)
            shards = P.CacheShards(rank_nodes, n_cached, wl["host_feat"], row_bytes, wl["feat_mask"],
                                   rank, world, dev, num_replicated=n_repl)
            hp.build_cache(rank_nodes, cache_pct, wl["host_feat"], row_bytes, wl["feat_mask"], num_shards=world,
                           shard_id=rank, peer_ptrs=shards.ptrs, fill_local=False, num_replicated=n_repl,
                           replica_ptr=shards.replica_ptr)
        else:
            hp.build_cache(rank_nodes, cache_pct, wl["host_feat"], row_bytes, wl["feat_mask"])
        torch.cuda.synchronize()
        cache_s = time.time() - t0
        hist = torch.zeros((Ksteps, hp.L, 3), dtype=torch.int32, device=dev)
        n_groups = (Ksteps + G - 1) // G
        gev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(n_groups)]
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(Ksteps)]
        sampled = [torch.cuda.Event() for _ in range(2)]
        gathered = [torch.cuda.Event() for _ in range(2)]
        main = torch.cuda.current_stream()

        def one_group(gi, k0, nb, key, timed):
            """steps k0 .. k0+nb-1 (timed: index of the first one in the timed region, or None)"""
            grp = gi % 2
            batches = []
            for j in range(nb):
                sd, n = seeds_of(k0 + j)
                batches.append((sd, n, key + j, grp * G + j))
            with torch.cuda.stream(s_streams[grp]):
                s_streams[grp].wait_event(gathered[grp])       # the group's previous batches have been extracted
                if timed is not None:
                    gev[timed // G][0].record()
                if nb == 1:
                    hp.sample(*batches[0][:3], slot=batches[0][3])
                else:
                    hp.sample_multi(batches)
                if timed is not None:
                    gev[timed // G][1].record()
                sampled[grp].record()
            with torch.cuda.stream(x_stream):
                x_stream.wait_event(sampled[grp])
                for j, (sd, n, _, slot) in enumerate(batches):
                    e = ev[timed + j] if timed is not None else None
                    if e:
                        e[0].record()
                    hp.gather(slot)
                    if e:
                        e[1].record()
                    hp.gather_labels(sd, n)
                    if e:
                        hist[timed + j].copy_(hp.slots[slot].counts)
                        e[2].record()
                gathered[grp].record()

        def run(first_step, nsteps, key, timed):
            gi = 0
            for k0 in range(0, nsteps, G):
                nb = min(G, nsteps - k0)
                one_group(gi, first_step + k0, nb, key + k0, k0 if timed else None)
                gi += 1

        for st in s_streams + [x_stream]:
            st.wait_stream(main)
        run(0, max(W, 2 * G), key0, False)       # warm-up covers both slot groups
        Wd = max(W, 2 * G)
        torch.cuda.synchronize()
        hp.stats.zero_()
        hp.remote.zero_()
        torch.cuda.synchronize()
        clocks = ClockSampler(local)
        clocks.start()
        if world > 1:
            dist.barrier()
        launches0 = K.launch_count()
        clocks.begin()
        torch.cuda.synchronize()
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        if profile:
            torch.cuda.profiler.start()       # ncu --profile-from-start off captures exactly the timed region
        t_start.record()
        for st in s_streams + [x_stream]:
            st.wait_stream(main)
        th0 = time.perf_counter()
        run(Wd, Ksteps, key0 + Wd, True)
        host_enqueue_s = time.perf_counter() - th0
        for st in s_streams + [x_stream]:
            main.wait_stream(st)
        t_end.record()
        torch.cuda.synchronize()
        if profile:
            torch.cuda.profiler.stop()
        if world > 1:
            dist.barrier()
        r = dict(cache_s=cache_s, launches=K.launch_count() - launches0, clk=clocks.stop(),
                 ms_total=t_start.elapsed_time(t_end), host_enqueue_ms=host_enqueue_s * 1e3)
        h = hist.cpu().numpy().astype("int64")
        r["edges"] = int(h[:, :, 1].sum())
        r["n_in_total"] = int(h[:, 0, 2].sum())       # input_nodes of every step (num_src of layer 0)
        r["sample_ms"] = sum(e[0].elapsed_time(e[1]) for e in gev)   # the slot groups' stream time of the sampling chain
        r["gather_ms"] = sum(e[0].elapsed_time(e[1]) for e in ev)    # the gather kernel alone, on its stream
        r["extract_ms"] = sum(e[0].elapsed_time(e[2]) for e in ev)   # gather + label gather
        r["hits"], r["misses"] = [int(x) for x in hp.stats.tolist()]
        r["remote"] = int(hp.remote.item())           # rows read from peer shards over NVLink (counted by the kernel)
        if shards is not None:
            r["shard_bytes"] = shards.nbytes
            r["replica_bytes"] = shards.replica_bytes
            hp.cache_table = hp.shard_ptrs = None
            shards.close()
        return r

    Ksteps, W = args.steps, max(3, args.warmup)
    # reference-like regime first (25 % cache, misses over the host link) ...
    if args.only_partition:
        args.no_cache25 = args.no_e2e = args.no_cpu_baseline = True
    r25 = measure(0.25, min(Ksteps, steps_per_epoch), W, 2_000_000) \
        if args.cache_pct != 0.25 and not args.no_cache25 else None
    # ... then the headline regime
    r = measure(args.cache_pct, Ksteps, W, 0, profile=True)
    ms_total, edges, n_in_total = r["ms_total"], r["edges"], r["n_in_total"]
    sample_ms, gather_ms, hits, misses = r["sample_ms"], r["gather_ms"], r["hits"], r["misses"]
    launches, clk, cache_s = r["launches"], r["clk"], r["cache_s"]

    # the dominant kernel timed alone (no sampling in flight), same cache, the last batch's input nodes
    torch.cuda.synchronize()
    ga0, ga1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_alone = int(hp.slots[0].num_items.item())
    for _ in range(3):
        hp.gather(0)
    ga0.record()
    for _ in range(20):
        hp.gather(0)
    ga1.record()
    torch.cuda.synchronize()
    gather_alone_ms = ga0.elapsed_time(ga1) / 20
    # serialised shares (one batch at a time on ONE stream: sample chain -> gather -> labels), the quantity an ncu
    # launch list of this command shows (its replay serialises the kernels); never used for `value`
    serial = None
    try:
        n_ser = 40
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_ser)]
        for k in range(n_ser + 3):
            sd, n = seeds_of(k)
            e = evs[k - 3] if k >= 3 else None
            if e:
                e[0].record()
            hp.sample(sd, n, 5_000_000 + k, slot=0)
            if e:
                e[1].record()
            hp.gather(0)
            if e:
                e[2].record()
            hp.gather_labels(sd, n)
            if e:
                e[3].record()
        torch.cuda.synchronize()
        s_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / n_ser
        g_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / n_ser
        l_ms = sum(e[2].elapsed_time(e[3]) for e in evs) / n_ser
        serial = {"sample_chain_ms": s_ms, "gather_ms": g_ms, "labels_ms": l_ms,
                  "gather_share": round(g_ms / (s_ms + g_ms + l_ms), 3),
                  "note": "one batch at a time on one stream; compare THIS share with the ncu launch list "
                          "(profiles/*_launches_timed_region.txt), whose replay serialises the kernels"}
    except Exception as ex:      # diagnostic only: never fail the bench line for it
        serial = {"error": repr(ex)[:200]}
    ms_total, edges_all = aggregate(ms_total, edges, dev)

    # ---- N > 1: the same workload with the cache PARTITIONED over the ranks' GPUs (north_star).
    # hybrid: every GPU keeps the hottest --replicate-pct of the vertices, the tail of the cache is striped over the
    # GPUs and read by NVLink peer loads inside the gather kernel; striped: no replicated head (round 1's layout).
    part = part_striped = part_cap = None
    if world > 1 and not args.no_partition:
        def part_leg(replicate_pct, key0):
            hp.cache = None
            torch.cuda.empty_cache()
            kp = Ksteps
            rp = measure(args.cache_pct, kp, W, key0, partition=True, replicate_pct=replicate_pct)
            p_ms, p_edges = aggregate(rp["ms_total"], rp["edges"], dev)
            remote = rp["remote"] * row_bytes
            return {"edges_per_s": p_edges / (p_ms * 1e-3), "ms_per_step": p_ms / kp, "steps": kp,
                    "replicate_pct": replicate_pct, "gather_ms_per_step": rp["gather_ms"] / kp,
                    "shard_GB_per_gpu": rp["shard_bytes"] / 1e9, "replica_GB_per_gpu": rp["replica_bytes"] / 1e9,
                    "remote_row_fraction": rp["remote"] / max(1, rp["hits"] + rp["misses"]),
                    "nvlink_peer_GBps_per_gpu": remote / (rp["gather_ms"] * 1e-3) / 1e9 if rp["gather_ms"] else None,
                    "nvlink_peak_GBps": 900.0, "cache_hit_rate": rp["hits"] / max(1, rp["hits"] + rp["misses"]),
                    "extract_GBps_per_gpu": rp["n_in_total"] * row_bytes / (rp["gather_ms"] * 1e-3) / 1e9,
                    "edges": p_edges, "ms_total": p_ms, "raw": rp}
        if args.only_partition:
            # diagnosis sweep: replicated head size x how peer rows are fetched (bulk engine / warp loads)
            os.environ["FGNN_TUNING_DYNAMIC"] = "1"
            diag = []
            for rp, dfr in ((0.25, 1), (0.25, 0), (0.1, 1), (0.1, 0), (0.0, 1)):
                os.environ["FGNN_GATHER_DEFER"] = str(dfr)
                d = part_leg(rp, 5_000_000 + int(rp * 100) * 10 + dfr)
                diag.append({"replicate_pct": rp, "deferred_peer_pass": dfr, "ms_per_step": round(d["ms_per_step"], 4),
                             "gather_ms_per_step": round(d["gather_ms_per_step"], 4),
                             "remote_row_fraction": round(d["remote_row_fraction"], 4),
                             "nvlink_GBps": round(d["nvlink_peer_GBps_per_gpu"] or 0, 1)})
                if rank == 0:
                    print("PARTITION_DIAG " + json.dumps(diag[-1]), file=sys.stderr, flush=True)
            os.environ.pop("FGNN_GATHER_DEFER", None)
        part = part_leg(args.replicate_pct, 3_000_000)
        part["note"] = ("cache partitioned over the %d GPUs: the hottest %.0f%% of the vertices (PreSC ranks) on every "
                        "GPU, the rest striped (slot %% N) and read by NVLink peer loads inside fgnn_k_gather_cached_layout; "
                        "remote rows counted by the kernel; population = PreSC ranking broadcast + each rank fills its "
                        "own stripe and replica" % (world, args.replicate_pct * 100))
        if args.replicate_pct > 0:
            part_striped = part_leg(0.0, 4_000_000)
            part_striped["note"] = ("no replicated head: every cached row striped over the GPUs (round 1's layout); most "
                                    "remote rows are HOT rows here, which the requesting GPU's L2 and the owner's L2 absorb")
        if args.replicate_pct > 0.25:
            part_cap = part_leg(0.25, 4_500_000)
            part_cap["note"] = ("capacity-oriented point: only the hottest 25% replicated; the ~3% of rows that are then "
                                "read from peers are COLD rows, which the fabric serves at a few tens of GB/s per GPU while "
                                "every GPU's HBM is saturated by its own gather (profiles/r2_partition_diag_n4.txt)")

    replicated = None
    if part is not None:
        # N > 1: the headline is the PARTITIONED arm (north_star); the replicated-cache arm moves to extra
        replicated = {"note": "every GPU holds the whole cache (the 57 GB table fits one B200): N independent replicas, "
                              "no NVLink traffic", "edges_per_s": edges_all / (ms_total * 1e-3),
                      "ms_per_step": ms_total / Ksteps, "gather_ms_per_step": gather_ms / Ksteps}
        hr = part.pop("raw")
        ms_total, edges_all = part.pop("ms_total"), part.pop("edges")
        edges, n_in_total = hr["edges"], hr["n_in_total"]
        sample_ms, gather_ms, hits, misses = hr["sample_ms"], hr["gather_ms"], hr["hits"], hr["misses"]
        launches, clk, cache_s, r = hr["launches"], hr["clk"], hr["cache_s"], hr
        for extra_leg in (part_striped, part_cap):
            if extra_leg is not None:
                for k in ("raw", "ms_total", "edges"):
                    extra_leg.pop(k)

    # ---- roofline of the dominant kernel: the fused cache-aware feature gather -----------
    peak, peak_kind = peaks()
    alg_bytes = n_in_total * (4 + 2 * row_bytes)              # B_ext = N_in*(4 + 2*D*4)
    achieved = alg_bytes / (gather_ms * 1e-3) / 1e9 if gather_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "gather_bulk_kernel<6,16> (fgnn_k_gather_cached: cp.async.bulk ring, 16 warps x 6 stages)",
                "achieved": round(achieved, 1), "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": round(achieved / peak, 4),
                "traffic": (ncu_traffic() or {}).get("dram_bytes_per_launch"), "traffic_capture": ncu_traffic(),
                "bytes_per_launch": alg_bytes // max(1, Ksteps), "avg_launch_ms": gather_ms / max(1, Ksteps),
                "share_of_step": round(gather_ms / ms_total, 3),
                "share_note": "gather stream time / wall time of the OVERLAPPED loop (sampling of later batches runs "
                              "next to it, so the per-stream times add up to more than the wall time)",
                "serialised": serial,
                "timing": "CUDA events on the extraction stream around every gather launch of the timed region; "
                          "super-batches of %d mini-batches are sampled concurrently on two other streams and share "
                          "HBM with it" % G,
                "alone": {"avg_launch_ms": gather_alone_ms, "bytes_per_launch": n_alone * (4 + 2 * row_bytes),
                          "achieved": round(n_alone * (4 + 2 * row_bytes) / (gather_alone_ms * 1e-3) / 1e9, 1),
                          "frac": round(n_alone * (4 + 2 * row_bytes) / (gather_alone_ms * 1e-3) / 1e9 / peak, 4),
                          "note": "same kernel, 20 back-to-back launches with nothing else on the GPU"}}

    out = {
        "metric": METRIC, "value": edges_all / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": Ksteps, "warmup": W, "ms_per_step": ms_total / Ksteps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": shared_config(args, V, wl["E"], D, int(wl["host_feat"].shape[0])),
        "notes": {"dtype_note": "u32 ids / hashes / counts; f32 feature rows and i64 labels are byte copies",
                  "batches_in_flight": len(hp.slots), "super_batch": G,
                  "cache_GB": hp.num_cached * row_bytes / 1e9,
                  "sharding": "seed mini-batches split across ranks (no data-path collective); topology replicated; "
                              "N = 1: whole cache in HBM; PreSC ranking: NCCL all-reduce + broadcast at init"},
        "clocks": clk, "gpu_launches": int(launches),
        "roofline": roofline,
        "extra": {"sample_only_edges_per_s": edges / (sample_ms * 1e-3) if sample_ms else None,
                  "extract_GBps": n_in_total * row_bytes / (gather_ms * 1e-3) / 1e9 if gather_ms else None,
                  "epoch_time_s_est": ms_total / Ksteps * steps_per_epoch / world * 1e-3,
                  "edges_per_step": edges / Ksteps, "input_nodes_per_step": n_in_total / Ksteps,
                  "cache_hit_rate": hits / max(1, hits + misses), "presc_s": presc_s, "cache_build_s": cache_s,
                  "graph_gen_s": wl["gen_s"], "sample_ms_per_step": sample_ms / Ksteps,
                  "gather_ms_per_step": gather_ms / Ksteps, "extract_ms_per_step": r["extract_ms"] / Ksteps,
                  "slots": len(hp.slots), "host_enqueue_ms_per_step": r["host_enqueue_ms"] / Ksteps,
                  "note": "sample_ms is the sampling chain's time on its own stream while other slots and the "
                          "gather run concurrently; value uses the wall time of the whole overlapped loop"},
    }
    if part is not None:
        out["extra"]["partitioned_cache"] = part
        out["extra"]["replicated_cache"] = replicated
        if part_striped is not None:
            out["extra"]["partitioned_cache_striped"] = part_striped
        if part_cap is not None:
            out["extra"]["partitioned_cache_replicate25"] = part_cap
        out["notes"]["sharding"] = ("seed mini-batches split across ranks (no data-path collective); topology replicated; "
                                     "feature cache PARTITIONED over the GPUs for `value` (hottest %.0f%% of the vertices on "
                                     "every GPU, the rest striped and read by NVLink peer loads); PreSC ranking: NCCL "
                                     "all-reduce + broadcast at init; extra.replicated_cache = a full replica per GPU"
                                     % (args.replicate_pct * 100))
    if r25 is not None:
        k25 = min(Ksteps, steps_per_epoch)
        out["extra"]["cache25"] = {
            "note": "same workload with the reference's 25 % cache: misses are read from pinned host memory (UVA)",
            "edges_per_s": r25["edges"] / (r25["ms_total"] * 1e-3), "ms_per_step": r25["ms_total"] / k25,
            "extract_ms_per_step": r25["extract_ms"] / k25,
            "cache_hit_rate": r25["hits"] / max(1, r25["hits"] + r25["misses"]),
            "miss_path_host_link_GBps": r25["misses"] * row_bytes / (r25["gather_ms"] * 1e-3) / 1e9,
            "extract_GBps": r25["n_in_total"] * row_bytes / (r25["gather_ms"] * 1e-3) / 1e9}

    # ---- e2e through the host runtime (samgraph_* C-ABI) with host buffers ------------------
    if not is_headline(args):
        out["notes"]["legs"] = ("non-headline sampler / fanout / workload: the CPU baseline is the reference's own code for "
                                "the uniform samplers and the oracle port for the others (see cpu_baseline.kind)")
    if not args.no_e2e:
        # release the device-resident leg's buffers first: the engine children build their own caches
        hp.cache = hp.feat_out = hp.cache_table = None
        del hp, rank_nodes
        torch.cuda.empty_cache()
        wl.update(tables)                # weighted samplers: the engine loads prob / alias / prefix tables from disk
        path = write_dataset_shm(args, wl, rank, world)
        try:
            out["e2e"] = run_e2e(args, wl, world, rank, dev, path)
            if world == 1 and rank == 0 and not args.no_cpu_baseline:
                # the reference's own CPU code on this box's cores (needs the graph: before it is released below)
                out["cpu_baseline"] = cpu_baseline(args, wl, steps=None, budget_s=20.0)
                args.no_cpu_baseline = True
            if not args.no_factored:
                # the FACTORED mode (FGNN: dedicated sampler GPUs + dedicated trainer GPUs, arch5) on the same N GPUs:
                # rank 0 forks S sampler and T trainer processes, the other ranks keep their GPUs idle meanwhile
                wl_small = {k: wl[k] for k in ("V", "E", "D", "C", "T")}
                for k in ("indptr", "indices", "label", "train") + tuple(tables):
                    wl.pop(k, None)
                tables.clear()
                torch.cuda.empty_cache()
                fact = run_factored(args, wl_small, world, rank, path)
                if fact:
                    out["e2e"]["factored"] = fact["e2e_factored"]
                    out["extra"]["epoch"] = fact["epoch"]
        finally:
            if world > 1:
                dist.barrier()
            if rank == 0 and not os.environ.get("FGNN_BENCH_KEEP_DATASET"):
                import shutil
                shutil.rmtree(path, ignore_errors=True)
    # ---- CPU baseline: the reference's own CPU code on this box's cores -----------------------
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, wl, steps=None, budget_s=20.0)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def write_dataset_shm(args, wl, rank, world):
    """Persist the generated graph in the reference's on-disk format (meta.txt + *.bin, engine.cc:73-264) on
    tmpfs so the C++ engine can load it like a real dataset.  feat.bin is omitted: SAMGRAPH_EMPTY_FEAT=k."""
    import numpy as np
    import torch.distributed as dist
    path = "/dev/shm/fgnn_bench_%s" % args.workload
    if rank == 0:
        os.makedirs(path, exist_ok=True)
        T = wl["T"]
        with open(os.path.join(path, "meta.txt"), "w") as f:
            f.write("NUM_NODE %d\nNUM_EDGE %d\nFEAT_DIM %d\nNUM_CLASS %d\nNUM_TRAIN_SET %d\nNUM_VALID_SET %d\nNUM_TEST_SET %d\n"
                    % (wl["V"], wl["E"], wl["D"], wl["C"], T, 16, 16))
        wl["indptr"].cpu().numpy().tofile(os.path.join(path, "indptr.bin"))
        wl["indices"].cpu().numpy().tofile(os.path.join(path, "indices.bin"))
        wl["label"].cpu().numpy().tofile(os.path.join(path, "label.bin"))
        wl["train"].cpu().numpy().tofile(os.path.join(path, "train_set.bin"))
        np.arange(16, dtype=np.uint32).tofile(os.path.join(path, "test_set.bin"))
        np.arange(16, dtype=np.uint32).tofile(os.path.join(path, "valid_set.bin"))
        for key, fname in (("prob_table", "prob_table.bin"), ("alias_table", "alias_table.bin"),
                           ("prefix_table", "prob_prefix_table.bin")):
            if wl.get(key) is not None:
                wl[key].cpu().numpy().tofile(os.path.join(path, fname))
    if world > 1:
        dist.barrier()
    return path


E2E_CACHE_PCT = 0.25   # reference-like regime for the end-to-end leg: 75 % of the feature table stays in host memory


def factored_split(n_gpus):
    """(samplers, trainers, single_gpu) of the factored legs on n_gpus GPUs: the reference's placements
    (multi_gpu/common_config.py:182-185; the paper runs 2 samplers + 6 trainers on 8 GPUs)."""
    if n_gpus <= 1:
        return 1, 1, True
    s = max(1, n_gpus // 4)
    return s, n_gpus - s, False


def run_factored(args, wl, world, rank, path):
    """FGNN's factored mode through the public API, on all `world` GPUs of the box: S sampler processes and T
    trainer processes forked by examples/train_graphsage_multi_gpu.py (the DGL-free mirror of the reference's
    multi_gpu/train_graphsage.py), cache partitioned over the trainers (hybrid), tasks through the device ring.
      e2e_factored : --no-train, pipeline mode: sampled edges/s delivered to the trainers, host clock, labels read back
      epoch        : with the GraphSAGE model (mean aggregation as CSR SpMM, Adam, DDP over NCCL between trainers):
                     measured epoch time and the kLogEpoch{Sample,Copy,Train}Time the scripts print
    Runs on rank 0 only; the other ranks wait on a CPU-side (gloo) barrier so that their GPUs stay idle."""
    import torch.distributed as dist
    S, T, single = factored_split(world)
    grp = dist.new_group(backend="gloo") if world > 1 else None
    res = None
    if rank == 0:
        env = dict(os.environ, SAMGRAPH_EMPTY_FEAT=str(args.empty_feat), SAMGRAPH_LOG_LEVEL="error")
        for k in list(env):
            # the children are NOT torchrun workers: with TORCHELASTIC_USE_AGENT_STORE left set, their own
            # init_process_group(tcp://...) connects as a client to an agent store that does not exist and hangs
            if k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "GROUP_RANK", "LOCAL_WORLD_SIZE",
                     "ROLE_RANK", "ROLE_WORLD_SIZE", "GROUP_WORLD_SIZE", "ROLE_NAME") or k.startswith("TORCHELASTIC_"):
                env.pop(k, None)
        base = [sys.executable, os.path.join(ROOT, "examples", "train_graphsage_multi_gpu.py"), "--dataset-path", path,
                "--num-sample-worker", str(S), "--num-train-worker", str(T), "--sample-type", args.sample_type,
                "--cache-percentage", str(E2E_CACHE_PCT), "--pipeline", "--json", "--fanout"] + \
               [str(f) for f in fanouts_of(args)]
        if single:
            base.append("--single-gpu")

        def one(extra, tag):
            r = subprocess.run(base + extra, capture_output=True, text=True, env=env, timeout=1500)
            for line in r.stdout.splitlines():
                if line.startswith("FACTORED_JSON "):
                    return json.loads(line[len("FACTORED_JSON "):])
            err = r.stderr.strip()
            sys.stderr.write("---- %s leg stderr (tail) ----\n%s\n" % (tag, err[-6000:]))     # for the run's log
            # one readable line for the JSON: the example's own verdict first ("timeout after ...", "a worker failed"),
            # then exception / abort lines; the interleaved per-process stack dumps stay in the log above
            lines = err.splitlines()
            keep = [l.strip() for l in lines if l.startswith("train_graphsage_multi_gpu:")]
            keep += [l.strip() for l in lines if ("Error" in l or "what():" in l or "CHECK" in l) and len(l) < 300][:6]
            return {"error": "%s leg failed: %s" % (tag, " | ".join(dict.fromkeys(keep))[:1200] or err[-400:])}
        # the consumer of the epoch leg is the reference's model for the configuration: GraphSAGE (headline, weighted),
        # GCN for the three-layer k-hop config (#4), PinSAGE on the random-walk blocks (#3)
        model = "pinsage" if args.sample_type == "random_walk" else ("gcn" if len(fanouts_of(args)) == 3 else "graphsage")
        margs = [] if model == "graphsage" else ["--model", model]
        f = one(["--no-train", "--num-epoch", "4", "--timeout", "120"], "e2e_factored")
        e = one(["--num-epoch", "3", "--timeout", "120"] + margs, "epoch")
        ddp_error = None
        if "error" in e and T > 1:
            # the measured epoch must not depend on the gradient all-reduce coming up: retry without DDP and say so
            ddp_error = e["error"]
            e = one(["--num-epoch", "3", "--no-ddp", "--timeout", "120"] + margs, "epoch")
        res = {}
        if "error" in f:
            res["e2e_factored"] = f
        else:
            timed = f["epochs"][1:]
            res["e2e_factored"] = {
                "value": f["edges_per_s"], "unit": UNIT, "samplers": S, "trainers": T, "single_gpu": single,
                "steps": f["steps_timed"], "ms_per_step": 1e3 * sum(x["wall_s"] for x in timed) / max(1, f["steps_timed"]),
                "cache_percentage": E2E_CACHE_PCT, "epoch_wall_s": [round(x["wall_s"], 5) for x in f["epochs"]],
                "per_trainer_wall_s_last_epoch": f["epochs"][-1]["per_trainer_wall_s"],
                "h2d_bytes_per_step": int(sum(x["miss_bytes"] for x in timed) / max(1, f["steps_timed"])),
                "d2h_bytes_per_step": 8 * BATCH, "init_s": f["init_s"],
                "api": "samgraph.torch arch5: data_init + fork, sample_init / train_init, sample_once on the sampler "
                       "GPUs, extract_start + get_next_batch + get_graph_* on the trainer GPUs",
                "note": "whole job on the box's GPUs: S sampler GPUs feed T trainer GPUs through the device ring "
                        "(payload in trainer HBM, NVLink peer copies); feature cache %d%% of the vertices, partitioned over "
                        "the trainers (hottest ranks replicated, tail striped, NVLink peer loads), misses from pinned "
                        "host memory; first epoch is warm-up" % int(E2E_CACHE_PCT * 100)}
        if "error" in e:
            res["epoch"] = e
        else:
            timed = e["epochs"][1:]
            k = len(timed)
            res["epoch"] = {
                "epoch_time_s": e["avg_epoch_s"], "samplers": S, "trainers": T, "single_gpu": single,
                "steps_per_epoch": e["epochs"][0]["steps"], "epochs_timed": k,
                "kLogEpochSampleTime_s": sum(x["sample_s"] for x in timed) / k,
                "kLogEpochCopyTime_s": sum(x["copy_s"] for x in timed) / k,
                "kLogEpochConvertTime_s": sum(x["convert_s"] for x in timed) / k,
                "kLogEpochTrainTime_s": sum(x["train_s"] for x in timed) / k,
                "epoch_wall_s": [round(x["wall_s"], 5) for x in e["epochs"]], "loss": e["loss"],
                "ddp": ddp_error is None and T > 1, "ddp_error": ddp_error,
                "model": "%s %d layers, hidden 256, aggregation as CSR SpMM on the CSC hand-off, Adam, "
                         "DDP (NCCL) between the trainers; examples/train_graphsage_multi_gpu.py --pipeline"
                         % ({"graphsage": "GraphSAGE (mean)", "gcn": "GCN (GraphConv norm=both)",
                             "pinsage": "PinSAGE (WeightedSAGEConv)"}[model], len(fanouts_of(args))),
                "note": "measured, not extrapolated: whole epochs of the papers100M-shaped train set through samgraph.torch "
                        "with the model as consumer; max over trainers; first epoch dropped like the scripts do"}
    if world > 1:
        dist.barrier(group=grp)
    return res


def run_e2e(args, wl, world, rank, dev, path):
    """Same metric through the public API: a child process drives the C++ engine (samgraph.torch over the
    samgraph_* C-ABI) on the dataset loaded from disk; see tools/e2e_runtime.py.  Two regimes:
    the headline one is the reference's situation (feature table in pinned HOST memory, PreSC cache 25 %, miss
    rows cross the host link inside every step); the all-in-HBM regime of `value` is reported next to it."""
    import torch
    import torch.distributed as dist
    # The end-to-end leg is always long enough to be trusted: at least one papers100M epoch
    # (151 steps) timed, after a warm-up that fills the pipeline (slot groups, max_sampling_jobs, the extractor's
    # depth) and grows the engine's device pools to their steady state.
    Ksteps, W = max(args.steps, 151), max(32, args.warmup)
    env = dict(os.environ, SAMGRAPH_EMPTY_FEAT=str(args.empty_feat), SAMGRAPH_LOG_LEVEL="error")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)

    def one(cache_pct):
        cmd = [sys.executable, os.path.join(ROOT, "tools", "e2e_runtime.py"), path, str(Ksteps), str(W),
               str(cache_pct), dev, str(0x5EED0000 + rank), args.sample_type, ",".join(str(f) for f in fanouts_of(args))]
        if world > 1:
            dist.barrier()
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1500)
        res = None
        for line in r.stdout.splitlines():
            if line.startswith("E2E_JSON "):
                res = json.loads(line[len("E2E_JSON "):])
        if res is None:
            raise RuntimeError("e2e runtime leg failed:\n%s\n%s" % (r.stdout[-2000:], r.stderr[-2000:]))
        if world > 1:
            # whole-job number: all ranks run concurrently; time = max over ranks, edges = sum
            t = torch.tensor([res["ms_per_step"]], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e = torch.tensor([res["edges_per_step"]], dtype=torch.float64, device=dev)
            dist.all_reduce(e, op=dist.ReduceOp.SUM)
            per_rank = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(per_rank, torch.tensor([res["ms_per_step"]], dtype=torch.float64, device=dev))
            res["per_rank_ms_per_step"] = [round(float(x.item()), 4) for x in per_rank]   # a straggler shows here
            res["ms_per_step"] = float(t.item())
            res["value"] = float(e.item()) / (res["ms_per_step"] * 1e-3)
            dist.barrier()
        return res

    res = one(E2E_CACHE_PCT)
    if args.cache_pct != E2E_CACHE_PCT:
        hbm = one(args.cache_pct)
        res["all_in_hbm"] = {k: hbm[k] for k in ("value", "unit", "ms_per_step", "h2d_bytes_per_step",
                                                  "d2h_bytes_per_step", "cache_percentage", "init_s", "steps",
                                                  "step_wall_us_p50_p99_max") if k in hbm}
    return res


# ---------------------------------------------------------------------------
def cpu_baseline(args, wl, steps, budget_s):
    """The CPU path on this host's cores, on a bounded sample of the same workload.
    Uniform samplers (khop2 / khop0): the reference's OWN code (oracle/_ref: CPUSampleKHop2|0 + CPUHashTable2 +
    CPUExtract, unmodified translation units), kind "reference".  The reference has no CPU code for the other
    samplers (cpu_sampling_{khop1,weighted_khop,random_walk}.cc are empty stubs): those are timed on the
    oracle's restatement of the CUDA algorithms (oracle/fgnn_oracle.c: OpenMP samplers, serial ordered unique),
    kind "port"."""
    import numpy as np
    fan = fanouts_of(args)
    st = getattr(args, "sample_type", "khop2")
    cores = os.cpu_count() or 1
    t0 = time.time()
    indptr = wl["indptr"].cpu().numpy().view(np.uint32)
    indices = wl["indices"].cpu().numpy().view(np.uint32)          # CPUSampleKHop2 permutes rows in place
    label = wl["label"].cpu().numpy()
    train = wl["train"].cpu().numpy().view(np.uint32)
    feat = wl["host_feat"].numpy()
    mask = np.uint32(feat.shape[0] - 1)
    if st in ("khop2", "khop0"):
        from oracle.oracle import RefCPU, have_ref
        if not have_ref():
            return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
        ref = RefCPU()
        ref.set_threads(cores)
        ht = ref.hashtable(2, wl["V"])
        sample = ref.sample_khop2 if st == "khop2" else ref.sample_khop0
        kind = "reference"
        what = "CPUSample%s + CPUHashTable2 Populate/MapNodes/MapEdges + CPUExtract" % ("KHop2" if st == "khop2" else "KHop0")

        def one_batch(seeds, key):
            ht.reset()
            ht.populate(seeds)
            cur, e_step = seeds, 0
            for i in range(len(fan) - 1, -1, -1):                  # cpu_loops.cc:55-191
                s, d = sample(indptr, indices, cur, fan[i])
                ht.populate(d)
                cur = ht.map_nodes()
                ht.map_edges(s, d)
                e_step += len(s)
            return cur, e_step
        extract = ref.extract
    else:
        from oracle.oracle import Oracle, sample_batch_oracle
        o = Oracle()
        o.lib.fgo_set_threads(cores)
        graph = dict(indptr=indptr, indices=indices)
        if st.startswith("weighted"):
            if "prob_table" in wl or "prefix_table" in wl:         # our arm: the tables the GPU builders made
                for k_src, k_dst in (("prob_table", "prob_table"), ("alias_table", "alias_table"),
                                     ("prefix_table", "prob_prefix_table")):
                    if k_src in wl:
                        a = wl[k_src].cpu().numpy()
                        graph[k_dst] = a.view(np.uint32) if k_src == "alias_table" else a
            else:                                                  # reference arm on a box without our kernels
                w = np.random.default_rng(SEED_WEIGHTS).integers(1, 11, size=len(indices)).astype(np.float32)
                if st == "weighted_khop_prefix":
                    graph["prob_prefix_table"] = o.build_prefix_table(indptr, w)
                else:
                    graph["prob_table"], graph["alias_table"] = o.build_alias_table(indptr, indices, w)
        kind = "port"
        what = "oracle restatement of the CUDA %s sampler (OpenMP) + ordered unique / remap (serial) + row gather" % st

        def one_batch(seeds, key):
            b = sample_batch_oracle(o, graph, seeds, fan, st, 0x5EED, key, rw=RW if st == "random_walk" else None)
            return b["input_nodes"], sum(l["num_edge"] for l in b["layers"])
        extract = o.extract
    prep = time.time() - t0
    rng = np.random.default_rng(0)
    order = rng.permutation(len(train))
    done, edges, n_in, spent = 0, 0, 0, 0.0
    k = 0
    max_steps = steps if steps is not None else 60
    spe = (len(train) + BATCH - 1) // BATCH
    while done < max_steps:
        kk = k % spe
        seeds = train[order[kk * BATCH:(kk + 1) * BATCH]]
        k += 1
        t1 = time.time()
        cur, e_step = one_batch(seeds, k)
        f = extract(feat, cur & mask)                              # DoFeatureExtract (mock-masked ids)
        lb = extract(label, seeds)
        dt = time.time() - t1
        if k == 1:
            continue                                               # warm-up step
        done += 1
        edges += e_step
        n_in += len(cur)
        spent += dt
        if steps is None and spent > budget_s:
            break
    return {"value": edges / spent if spent else None, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d mini-batches (batch %d, %s fanout %s) of the same graph: %s (feature rows masked to the "
                      "2^k-row host table), %d OpenMP threads; %.1f s" % (done, BATCH, st, fan, what, cores, spent),
            "steps": done, "ms_per_step": spent / max(1, done) * 1e3,
            "extract_rows_per_step": n_in / max(1, done), "prep_s": prep}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    dev = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0")) if torch.cuda.is_available() else "cpu"
    wl = build_workload(args, dev)
    K, W = args.steps, max(3, args.warmup)
    steps = min(K, 40)                                # bounded: each step is one CPU mini-batch
    cb = cpu_baseline(args, wl, steps=steps, budget_s=1e9)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": cb["steps"], "warmup": 1,
           "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": DTYPE, "data": "synthetic",
           "config": shared_config(args, wl["V"], wl["E"], wl["D"], int(wl["host_feat"].shape[0])),
           "notes": {"reference_path": "the CPU path on the host cores (cpu_baseline.kind / .sample say which code): every "
                                       "feature row is read from host memory (no GPU, no cache)"},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
