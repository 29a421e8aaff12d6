"""Debug aid: does torch's CUDA generator recover after a FAILED graph capture if a trivial capture succeeds afterwards?"""
import torch

dev = torch.device("cuda")
x = torch.ones(8, device=dev)
s = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    g.capture_begin(capture_error_mode="thread_local")
    try:
        y = x * 2
        float(y[0].item())
    except Exception as e:
        print("body failed:", str(e).splitlines()[0][:100])
        try:
            g.capture_end()
        except Exception as e2:
            print("capture_end failed:", str(e2).splitlines()[0][:100])
torch.cuda.synchronize()
try:
    torch.randn(4, device=dev)
    print("randn ok without recovery")
except Exception as e:
    print("randn broken:", str(e).splitlines()[0][:100])
    g2 = torch.cuda.CUDAGraph()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g2.capture_begin(capture_error_mode="thread_local")
        z = x + 1
        g2.capture_end()
    torch.cuda.synchronize()
    try:
        print("after dummy capture:", torch.randn(4, device=dev).shape, "ok")
    except Exception as e3:
        print("still broken:", str(e3).splitlines()[0][:100])
