#!/usr/bin/env python
"""Peer-to-peer topology of the box: nvidia-smi topo -m, can_device_access_peer and the measured copy bandwidth
between every pair of visible GPUs (256 MB cudaMemcpyPeer, best of 3).  One process, no NCCL."""
import json
import subprocess
import sys

import torch


def main():
    n = torch.cuda.device_count()
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout)
    except Exception as e:
        print("nvidia-smi topo failed:", e)
    nbytes = 256 << 20
    bufs = [torch.empty(nbytes, dtype=torch.uint8, device="cuda:%d" % i) for i in range(n)]
    out = {"n": n, "access": [[bool(torch.cuda.can_device_access_peer(i, j)) if i != j else True for j in range(n)]
                              for i in range(n)], "GBps": [[0.0] * n for _ in range(n)]}
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            best = 0.0
            for _ in range(3):
                torch.cuda.synchronize(i)
                torch.cuda.synchronize(j)
                with torch.cuda.device(j):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    bufs[j].copy_(bufs[i], non_blocking=True)       # device j pulls from device i
                    e1.record()
                    e1.synchronize()
                    best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
            out["GBps"][i][j] = round(best, 1)
    for i in range(n):
        print("from cuda:%d -> " % i + "  ".join("%7.1f" % out["GBps"][i][j] for j in range(n)))
    print("P2P_JSON " + json.dumps(out))


if __name__ == "__main__":
    main()
