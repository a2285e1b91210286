#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Measures the output path on the GPU box: device-side text formatting (k_text), the whole
write_vtk_structured_points call to a tmpfs file, the reference's Python writer loops on a bounded
sample of the same field, and probe-line downloads.  Prints one JSON object.
Usage: python tools/output_probe.py [--n 512] [--out gpurun_out/output_probe.json]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from adi_thermal_fields_b200 import _capi, vtk_writer as vw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    n = a.n
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    T = 20.0 + 1380.0 * torch.rand((n, n, n), dtype=torch.float64, device=dev, generator=g)
    L = _capi.load()
    ctx = _capi.context(0)
    res = {"shape": [n, n, n], "values": n ** 3}
    kc = 32
    cap = int(L.adi_text_capacity(n * n * kc))
    buf = torch.empty(cap, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for fmt, name in ((0, "e6"), (1, "g6")):
        nb = C.c_ulonglong(0)
        for rep in range(2):                         # first pass warms up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            total = 0
            for k0 in range(0, n, kc):
                _capi.check(L.adi_text_format(ctx, T.data_ptr(), 0, n, n, n, k0, min(kc, n - k0), fmt, buf.data_ptr(),
                                              cap, C.byref(nb), st), "adi_text_format")
                total += nb.value
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        res[f"format_{name}"] = {"seconds": dt, "text_bytes": total, "Gvalues_per_s": n ** 3 / dt / 1e9,
                                 "GB_per_s_algorithmic": (8 * n ** 3 + total) / dt / 1e9}
        tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
        path = os.path.join(tmp, f"adi_probe_{name}.vtk")
        w = vw.write_vtk_structured_points if fmt == 0 else vw.write_vtk_structured_points_mm
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nbytes = w(path, T, 1e-3)
            dt = time.perf_counter() - t0
        res[f"file_{name}"] = {"seconds": dt, "bytes": nbytes, "Mvalues_per_s": n ** 3 / dt / 1e6,
                               "file_GB_per_s": nbytes / dt / 1e9, "dir": tmp}
        os.remove(path)
        # the reference's loops (oracle restatement) on a bounded sample of the same field
        # (vtk_writer.py:4-9 / waam_from_stl_v7_mm.py:203-206: one Python format call per value)
        sample = T[:64, :64, :32].cpu().numpy()
        t0 = time.perf_counter()
        if fmt == 0:
            flat = sample.reshape(-1, order="F")
            "".join(" ".join(f"{float(v):.6e}" for v in flat[i:i + 9]) + "\n" for i in range(0, flat.size, 9))
        else:
            "".join(" ".join(f"{float(sample[i, j, k]):.6g}" for i in range(sample.shape[0])) + "\n"
                    for k in range(sample.shape[2]) for j in range(sample.shape[1]))
        dt = time.perf_counter() - t0
        res[f"reference_python_{name}"] = {"sample_values": sample.size, "seconds": dt,
                                           "Mvalues_per_s": sample.size / dt / 1e6, "cores": 1}
    # probes: cost on the stepping stream of one z line and one x-z slice, vs a blocking .cpu()
    rec = vw.ProbeRecorder(nslots=8, slot_bytes=n * n * 8)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for label, ix in (("line_z", (n // 2, n // 2, slice(None))), ("slice_xz", (slice(None), n // 2, slice(None))),
                      ("slice_xy", (slice(None), slice(None), n // 2))):
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            tk = rec.record(T, ix)
            e1.record()
            t_host = time.perf_counter() - t0
            got = rec.fetch(tk)
            t_fetch = time.perf_counter() - t0
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            ref = T[ix].cpu().numpy()
            t_sync = time.perf_counter() - t1
        assert np.array_equal(got, ref)
        res[f"probe_{label}"] = {"bytes": int(got.nbytes), "record_host_us": t_host * 1e6,
                                 "stream_us": e0.elapsed_time(e1) * 1e3, "record_to_fetch_us": t_fetch * 1e6,
                                 "blocking_index_cpu_us": t_sync * 1e6}
    s = json.dumps(res)
    print(s)
    if a.out:
        with open(a.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
