import importlib, sys, torch
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("2023-tifs-istvt_b200"); ops = pkg.ops
x = [torch.randn(384, 149, 149, 32, device="cuda").to(torch.bfloat16) for _ in range(2)]
w = (torch.randn(64, 3, 3, 32, device="cuda") * 0.05).to(torch.bfloat16); b = torch.zeros(64, device="cuda")
for i in range(3): ops.conv3x3(x[i % 2], w, b)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10): ops.conv3x3(x[i % 2], w, b)
e1.record(); torch.cuda.synchronize(); print("conv2 ms", e0.elapsed_time(e1) / 10)
