"""PCIe ceiling of the box for the e2e number: pinned D2H / H2D alone and both directions at once, sized like one C2 batch
(819 MB of dB rows out, 320 MB of PCM in).  usage (GPU box): python profiles/pcie_peak.py > gpurun_out/pcie_peak.json"""
import json
import sys
import time

import torch

dev = torch.device("cuda", 0)
# optional: bytes out / bytes in of another step shape, e.g. the uint8-spectrum step: python profiles/pcie_peak.py 214240976 160000000
out_b, in_b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (819_200_000, 320_000_000)
d_out = torch.empty(out_b, dtype=torch.uint8, device=dev)
d_in = torch.empty(in_b, dtype=torch.uint8, device=dev)
h_out = torch.empty(out_b, dtype=torch.uint8, pin_memory=True)
h_in = torch.empty(in_b, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def d2h():
    with torch.cuda.stream(s1):
        h_out.copy_(d_out, non_blocking=True)


def h2d():
    with torch.cuda.stream(s2):
        d_in.copy_(h_in, non_blocking=True)


def both():
    d2h()
    h2d()


t_d2h, t_h2d, t_both = timed(d2h), timed(h2d), timed(both)
print(json.dumps({"d2h_gbs": out_b / t_d2h / 1e9, "h2d_gbs": in_b / t_h2d / 1e9, "d2h_ms_819MB": 1e3 * t_d2h,
                  "h2d_ms_320MB": 1e3 * t_h2d, "both_ms": 1e3 * t_both, "bytes_out": out_b, "bytes_in": in_b,
                  "both_aggregate_gbs": (out_b + in_b) / t_both / 1e9,
                  "both_d2h_gbs_effective": out_b / t_both / 1e9,
                  "note": "one C2 batch moves 320 MB in and 828 MB out; both_ms is the PCIe floor of an e2e step"}))
