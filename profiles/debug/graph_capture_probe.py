"""Debug aid: find the call that invalidates a CUDA-graph capture of the iteration (prints the capture status after every
C-ABI call and at line granularity inside the captured body).  usage: python profiles/debug/graph_capture_probe.py K D desired updater"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gmmvi_b200 import ops, rng  # noqa: E402

cudart = ctypes.CDLL("libcudart.so.12")


def status():
    st = ctypes.c_int(-1)
    rc = cudart.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    return rc, st.value          # status 0 none, 1 active, 2 invalidated


orig = ops._call
bad = []


def probe(name, *a, **k):
    orig(name, *a, **k)
    rc, st = status()
    if (st == 2 or rc != 0) and not bad:
        bad.append(name)
        print("capture invalidated at/before C call:", name, rc, st, flush=True)


ops._call = probe
from test_graph_gpu import _fixed  # noqa: E402

K, D, desired, updater = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
rng.set_seed(11)
g = _fixed(K, D, desired, updater, False)
g.enable_cuda_graph()


def tracer(frame, event, arg):
    if event == "line" and "gmmvi_b200" in frame.f_code.co_filename:
        rc, st = status()
        if (st == 2 or rc != 0) and not bad:
            bad.append((frame.f_code.co_filename, frame.f_lineno))
            print("capture invalidated before", frame.f_code.co_filename, frame.f_lineno, rc, st, flush=True)
    return tracer


g.train_iter()
sys.settrace(tracer)
try:
    g.train_iter()
except Exception as e:
    print("capture failed:", str(e).splitlines()[0])
sys.settrace(None)
print("bad:", bad)
