"""Device-resident batch API: torch CUDA tensors in, torch CUDA tensors out, no host copies.

torch is plumbing only (device memory, streams, torch.distributed); every kernel is launched
through the C ABI's `*_dev` entry points on torch's current stream.  Tensors are int64 with the
ABI's u64 word layout, one row per element (same shapes as the numpy API: (n, 13) G1Affine, ...).
"""
import torch

from . import _native as nat


def _check(t, w, name):
    if not (t.is_cuda and t.dtype == torch.int64 and t.dim() == 2 and t.shape[1] == w and t.is_contiguous()):
        raise ValueError("%s must be a contiguous CUDA int64 tensor of shape (n, %d)" % (name, w))


class DeviceEngine:
    def __init__(self, ctx=None, device=None):
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.ctx = ctx if ctx is not None else nat.Context(self.device.index or 0)
        self._scratch = {}

    def _stream(self):
        # torch's default stream has handle 0, which the C ABI reads as "use the context's own stream";
        # name the legacy default stream explicitly (cudaStreamLegacy == 0x1) so that the kernels are
        # ordered with torch's events and collectives.
        return torch.cuda.current_stream(self.device).cuda_stream or 1

    def _buf(self, key, nbytes):
        """Grow-only named scratch buffer, so the timed path never allocates."""
        t = self._scratch.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)
            self._scratch[key] = t
        return t

    # ---------------------------------------------------------------- pairing engine
    def pairing(self, p, q, out=None):
        _check(p, nat.W_G1A, "p"); _check(q, nat.W_G2A, "q")
        n = p.shape[0]
        if out is None:
            out = torch.empty((n, nat.W_FQ12), dtype=torch.int64, device=self.device)
        self.ctx.call_dev("bls_pairing_dev", p.data_ptr(), q.data_ptr(), out.data_ptr(), n, self._stream())
        return out

    def pairing_projective(self, p, q, out=None):
        """Engine::pairing on Jacobian rows (n,18),(n,36): into_affine fused in front of the pairing."""
        _check(p, nat.W_G1, "p"); _check(q, nat.W_G2, "q")
        n = p.shape[0]
        if out is None:
            out = torch.empty((n, nat.W_FQ12), dtype=torch.int64, device=self.device)
        self.ctx.call_dev("bls_pairing_projective_dev", p.data_ptr(), q.data_ptr(), out.data_ptr(), n, self._stream())
        return out

    def miller_loop_batch(self, p, q, out=None):
        _check(p, nat.W_G1A, "p"); _check(q, nat.W_G2A, "q")
        n = p.shape[0]
        if out is None:
            out = torch.empty((n, nat.W_FQ12), dtype=torch.int64, device=self.device)
        self.ctx.call_dev("bls_miller_loop_dev", p.data_ptr(), q.data_ptr(), out.data_ptr(), n, self._stream())
        return out

    def g2_prepare(self, q, out=None):
        _check(q, nat.W_G2A, "q")
        n = q.shape[0]
        if out is None:
            out = torch.empty((n, nat.W_G2P), dtype=torch.int64, device=self.device)
        self.ctx.call_dev("bls_g2_prepare_dev", q.data_ptr(), out.data_ptr(), n, self._stream())
        return out

    def miller_loop_prepared_batch(self, p, qp, out=None):
        _check(p, nat.W_G1A, "p"); _check(qp, nat.W_G2P, "qp")
        n = p.shape[0]
        if out is None:
            out = torch.empty((n, nat.W_FQ12), dtype=torch.int64, device=self.device)
        self.ctx.call_dev("bls_miller_loop_prepared_dev", p.data_ptr(), qp.data_ptr(), out.data_ptr(), n, self._stream())
        return out

    def pairing_shared_q(self, p, q1_prepared, out=None, final_exp=True):
        """e(P_i, Q) for every P_i against ONE prepared Q ((1, 2449) tensor): coefficients staged in shared memory by TMA."""
        _check(p, nat.W_G1A, "p"); _check(q1_prepared, nat.W_G2P, "q1_prepared")
        n = p.shape[0]
        if out is None:
            out = torch.empty((n, nat.W_FQ12), dtype=torch.int64, device=self.device)
        self.ctx.call_dev("bls_miller_loop_shared_q_dev", p.data_ptr(), q1_prepared.data_ptr(), out.data_ptr(), n, 1 if final_exp else 0, self._stream())
        return out

    def final_exponentiation(self, f, out=None):
        _check(f, nat.W_FQ12, "f")
        n = f.shape[0]
        if out is None:
            out = torch.empty_like(f)
        ok = torch.empty(n, dtype=torch.uint8, device=self.device)
        self.ctx.call_dev("bls_final_exponentiation_dev", f.data_ptr(), out.data_ptr(), ok.data_ptr(), n, self._stream())
        return out, ok

    def multi_miller_loop(self, p, q, out=None):
        """Product of the Miller values of all pairs of this device's shard -> (1, 72)."""
        _check(p, nat.W_G1A, "p"); _check(q, nat.W_G2A, "q")
        n = p.shape[0]
        if out is None:
            out = torch.empty((1, nat.W_FQ12), dtype=torch.int64, device=self.device)
        scr = self._buf("mm", self.ctx.multi_miller_scratch_bytes(n))
        self.ctx.call_dev("bls_multi_miller_loop_dev", p.data_ptr(), q.data_ptr(), n, out.data_ptr(), scr.data_ptr(), self._stream())
        return out

    def pairing_product(self, p, q, out=None):
        """final_exponentiation(miller_loop(all pairs of this device)) -> ((1, 72), is_some (1,) uint8)."""
        _check(p, nat.W_G1A, "p"); _check(q, nat.W_G2A, "q")
        n = p.shape[0]
        if out is None:
            out = torch.empty((1, nat.W_FQ12), dtype=torch.int64, device=self.device)
        ok = torch.empty(1, dtype=torch.uint8, device=self.device)
        scr = self._buf("mm", self.ctx.multi_miller_scratch_bytes(n))
        self.ctx.call_dev("bls_pairing_product_dev", p.data_ptr(), q.data_ptr(), n, out.data_ptr(), ok.data_ptr(), scr.data_ptr(), self._stream())
        return out, ok

    def fq12_product_tail(self, f, final_exp=False, out=None):
        """Product of a few Fq12 values (per-device partials) by one block, optionally followed by the single final
        exponentiation on the warp-cooperative engine -> ((1, 72), is_some)."""
        _check(f, nat.W_FQ12, "f")
        if out is None:
            out = torch.empty((1, nat.W_FQ12), dtype=torch.int64, device=self.device)
        ok = torch.ones(1, dtype=torch.uint8, device=self.device)
        self.ctx.call_dev("bls_fq12_product_tail_dev", f.data_ptr(), f.shape[0], out.data_ptr(), 1 if final_exp else 0, ok.data_ptr(), self._stream())
        return out, ok

    def fq12_pow(self, a, k, out=None):
        _check(a, nat.W_FQ12, "a"); _check(k, nat.W_FR, "k")
        if out is None:
            out = torch.empty_like(a)
        self.ctx.call_dev("bls_fq12_pow_dev", a.data_ptr(), k.data_ptr(), out.data_ptr(), a.shape[0], self._stream())
        return out

    def fq12_product(self, f, out=None):
        _check(f, nat.W_FQ12, "f")
        n = f.shape[0]
        if out is None:
            out = torch.empty((1, nat.W_FQ12), dtype=torch.int64, device=self.device)
        scr = self._buf("prod", self.ctx.fq12_product_scratch_bytes(n))
        self.ctx.call_dev("bls_fq12_product_dev", f.data_ptr(), n, out.data_ptr(), scr.data_ptr(), self._stream())
        return out

    # ---------------------------------------------------------------- curves
    def g1_wnaf_mul(self, bases, k, window=0, out=None):
        _check(bases, nat.W_G1, "bases"); _check(k, nat.W_FR, "k")
        if out is None:
            out = torch.empty_like(bases)
        self.ctx.call_dev("bls_g1_wnaf_mul_dev", bases.data_ptr(), k.data_ptr(), out.data_ptr(), bases.shape[0], window, self._stream())
        return out

    def g2_wnaf_mul(self, bases, k, window=0, out=None):
        _check(bases, nat.W_G2, "bases"); _check(k, nat.W_FR, "k")
        if out is None:
            out = torch.empty_like(bases)
        self.ctx.call_dev("bls_g2_wnaf_mul_dev", bases.data_ptr(), k.data_ptr(), out.data_ptr(), bases.shape[0], window, self._stream())
        return out

    def wnaf_table(self, base, window, g2=False):
        """Wnaf::base(g, n): the shared window table (2^(window-1), W) on the device."""
        w = nat.W_G2 if g2 else nat.W_G1
        _check(base, w, "base")
        table = torch.empty((1 << (window - 1), w), dtype=torch.int64, device=self.device)
        self.ctx.call_dev("bls_g2_wnaf_table_dev" if g2 else "bls_g1_wnaf_table_dev", base.data_ptr(), window, table.data_ptr(), self._stream())
        return table

    def wnaf_fixed_base(self, table, window, k, g2=False, out=None):
        """`.scalar(k_i)` for every scalar against one shared table."""
        w = nat.W_G2 if g2 else nat.W_G1
        _check(table, w, "table"); _check(k, nat.W_FR, "k")
        if table.shape[0] != 1 << (window - 1):
            raise ValueError("table has %d entries, window %d needs %d" % (table.shape[0], window, 1 << (window - 1)))
        if out is None:
            out = torch.empty((k.shape[0], w), dtype=torch.int64, device=self.device)
        self.ctx.call_dev("bls_g2_wnaf_fixed_base_dev" if g2 else "bls_g1_wnaf_fixed_base_dev", table.data_ptr(), window, k.data_ptr(), out.data_ptr(), k.shape[0], self._stream())
        return out

    def g1_batch_normalization_(self, v):
        _check(v, nat.W_G1, "v")
        scr = self._buf("bn", self.ctx.batch_normalization_scratch_bytes(1, v.shape[0]))
        self.ctx.call_dev("bls_g1_batch_normalization_dev", v.data_ptr(), v.shape[0], scr.data_ptr(), self._stream())
        return v

    def g2_batch_normalization_(self, v):
        _check(v, nat.W_G2, "v")
        scr = self._buf("bn", self.ctx.batch_normalization_scratch_bytes(2, v.shape[0]))
        self.ctx.call_dev("bls_g2_batch_normalization_dev", v.data_ptr(), v.shape[0], scr.data_ptr(), self._stream())
        return v

    # normalised Jacobian -> affine rows (pure data movement: x, y, infinity flag = (z == 0))
    @staticmethod
    def jacobian_to_affine_rows(v, coord_words):
        n = v.shape[0]
        out = torch.zeros((n, 2 * coord_words + 1), dtype=torch.int64, device=v.device)
        out[:, :2 * coord_words] = v[:, :2 * coord_words]
        out[:, 2 * coord_words] = (v[:, 2 * coord_words:] == 0).all(dim=1).to(torch.int64)
        return out
