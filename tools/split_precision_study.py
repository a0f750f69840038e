"""CPU numerics study for the round-2 GEMM operand format (no GPU needed): error of split-operand products against an
fp64 reference, for the current 3xTF32 split and for the candidate fp16 hi/lo split (kind::f16 MMAs run at twice the TF32
rate and move half the shared-memory bytes per K).

  a = hi + lo,  product ~ hi*hi' + hi*lo' + lo*hi'   (the lo*lo' term, ~2^-22 relative, is dropped in both schemes)
  tf32 :  hi = rn_tf32(a),           lo = rn_tf32(a - hi)
  fp16 :  hi = rn_fp16(a) (sat.),    lo = rn_fp16((a - hi) * 2^11), cross terms scaled back by 2^-11 in the epilogue

Accumulation is emulated in fp64 (the tensor core's fp32 accumulate adds the same error to both schemes).
Prints max / rms relative-to-scale errors for GEMM shapes of the model and activation ranges from 1e-4 to 6e4."""
import torch


def rn_tf32(t):
    return ((t.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)


def split_tf32(t):
    hi = rn_tf32(t)
    return hi, rn_tf32(t - hi), 1.0


def split_f16(t):
    hi = t.clamp(-65504.0, 65504.0).half().float()
    lo = ((t - hi) * 2048.0).clamp(-65504.0, 65504.0).half().float()
    return hi, lo, 1.0 / 2048.0


def split_product(a, b, split):
    ah, al, sa = split(a)
    bh, bl, sb = split(b)
    d = lambda x, y: x.double() @ y.double()
    return d(ah, bh) + d(ah, bl) * sb + d(al, bh) * sa


def study(M, K, N, a_scale, seed=0):
    g = torch.Generator().manual_seed(seed)
    # activations: post-ReLU-like magnitudes spread over two decades around a_scale, 30 % zeros
    a = torch.randn(M, K, generator=g).abs() * a_scale * torch.exp(torch.randn(M, K, generator=g) * 1.5)
    a = a * (torch.rand(M, K, generator=g) > 0.3)
    w = torch.randn(K, N, generator=g) / K ** 0.5
    ref = a.double() @ w.double()
    scale = (a.double().abs() @ w.double().abs()).clamp_min(1e-300)          # sum |a_k w_k|: the natural error scale
    out = {}
    for name, sp in (("1xTF32", None), ("3xTF32", split_tf32), ("fp16 hi/lo", split_f16)):
        if sp is None:
            got = rn_tf32(a).double() @ rn_tf32(w).double()
        else:
            got = split_product(a, w, sp)
        e = ((got - ref).abs() / scale)
        out[name] = (e.max().item(), e.pow(2).mean().sqrt().item())
    return out, a.abs().max().item()


if __name__ == "__main__":
    print(f"{'shape (MxKxN)':>22s} {'a_scale':>9s} {'max|a|':>10s} | " + " | ".join(f"{n:>22s}" for n in ("1xTF32", "3xTF32", "fp16 hi/lo")))
    for (M, K, N) in ((256, 3456, 128), (256, 1152, 128), (256, 128, 384), (128, 18816, 512)):
        for a_scale in (1e-4, 1e-2, 1.0, 1e2, 2e3):
            res, amax = study(M, K, N, a_scale)
            print(f"{f'{M}x{K}x{N}':>22s} {a_scale:9.0e} {amax:10.3g} | " +
                  " | ".join(f"max {res[n][0]:.1e} rms {res[n][1]:.1e}" for n in ("1xTF32", "3xTF32", "fp16 hi/lo")))
