"""Diagnostic only (not product code): does a user-compiled tensor-mode TMA load run on this box at all?
Triton's own descriptor path is used as an independent witness next to tools/scratch/tma_min.cu; its PTX / SASS are
dumped to gpurun_out/ for a side-by-side with the hand-written probe."""
import os, sys
import torch
import triton
import triton.language as tl

try:
    from triton.tools.tensor_descriptor import TensorDescriptor
except Exception as e:  # noqa: BLE001
    print("no TensorDescriptor:", e)
    sys.exit(0)


@triton.jit
def k(desc, out_ptr, BM: tl.constexpr, BN: tl.constexpr):
    t = desc.load([4, 8])
    offs = tl.arange(0, BM)[:, None] * BN + tl.arange(0, BN)[None, :]
    tl.store(out_ptr + offs, t)


x = torch.arange(64 * 168, dtype=torch.float32, device="cuda").reshape(64, 168)
desc = TensorDescriptor.from_tensor(x, [16, 32])
out = torch.empty(16 * 32, device="cuda")
try:
    h = k[(1,)](desc, out, 16, 32)
    torch.cuda.synchronize()
    ok = torch.equal(out.reshape(16, 32), x[4:20, 8:40])
    print("triton TMA load ran, correct =", ok)
    os.makedirs("gpurun_out", exist_ok=True)
    for ext in ("ptx", "ttgir"):
        if ext in h.asm:
            open("gpurun_out/triton_tma." + ext, "w").write(h.asm[ext])
    ptx = h.asm.get("ptx", "")
    print([l.strip() for l in ptx.splitlines() if "cp.async.bulk" in l or "mbarrier" in l or ".target" in l or ".version" in l][:12])
    if "cubin" in h.asm:
        open("gpurun_out/triton_tma.cubin", "wb").write(h.asm["cubin"])
except Exception as e:  # noqa: BLE001
    print("triton TMA FAILED:", type(e).__name__, str(e)[:300])
