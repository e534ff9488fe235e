"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GP-prior KL hot path of ethanev/GP-VAE.

A float64 restatement (torch CPU, autograd for the backward) of the reference's algorithm.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (gp-vae_b200/gpkl) never does and fails loudly without its CUDA library.

Parity pinning: tests/test_oracle_golden.py checks this file against fixtures produced by running
the reference's OWN functions (unmodified, under oracle/tf1_stub.py) -- tests/golden/*.npz, made by
oracle/gen_golden.py.  The reference has no tests or golden vectors of its own.
Cauchy kernel and the bidiagonal posterior (V3) do not exist in the reference: for those this file
is the only definition -> "parity unpinned by reference".

Reference lines followed (relative to /root/reference/src/Models/):
  kernel build      Full_GP_VAE_dynamic_time.py:149-164   K = (1-noise)*exp(-d^2/(2 l^2)) + noise*I, float32
  sample            Full_GP_VAE_dynamic_time.py:165-171, :174-195   z_s = m + chol(K_q) eps_s
  pair ordering     Full_GP_VAE_dynamic_time.py:216-217, :231-240   p = b*D + d
  KL (V1)           Full_GP_VAE_dynamic_time.py:242-260   float64: LU inverse, two logdets, matmul+trace
  prior lengthscale Full_GP_VAE_fixed_for_MovMnist.py:88-100  (trainable l_p)
  KL (V2, diag q)   VAE_GPprior_diag_cov.py:73-119 ; sample :64-71 ; numpy kernel :153-165 (== noise 0)
  ragged batches    DataHandler.py:111-156  (times zero padded to T_max, lengths[B])
"""
from __future__ import annotations

import numpy as np
import torch

RBF, CAUCHY = "rbf", "cauchy"


def kernel_matrix(t: torch.Tensor, ell: torch.Tensor, kernel: str = RBF, noise: float = 1e-3,
                  build_dtype=torch.float32) -> torch.Tensor:
    """K[..., T, T] for times t[..., T] and lengthscales ell[...] (broadcast against t's batch dims).

    Follows tf_kernel (Full_GP_VAE_dynamic_time.py:154-164): built in float32 like the reference,
    returned in float64.  Cauchy (extension): (1-noise)/(1+d^2/l^2) + noise*I.
    """
    t = t.to(build_dtype)
    ell = ell.to(build_dtype)
    diff = t.unsqueeze(-1) - t.unsqueeze(-2)
    l2 = (ell * ell).unsqueeze(-1).unsqueeze(-1)
    if kernel == RBF:
        k = torch.exp(-(diff * diff) / (2.0 * l2))
    elif kernel == CAUCHY:
        k = 1.0 / (1.0 + (diff * diff) / l2)
    else:
        raise ValueError(kernel)
    T = t.shape[-1]
    K = (1.0 - noise) * k + noise * torch.eye(T, dtype=build_dtype)
    return K.to(torch.float64)


def _logdet(K):  # tf.linalg.logdet == 2*sum(log(diag(chol)))  (Full_GP_VAE_dynamic_time.py:251-252)
    L = torch.linalg.cholesky(K)
    return 2.0 * torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)


def kl_full_gp(m, Kq, Kp, formulation="reference"):
    """gp_kl_div (Full_GP_VAE_dynamic_time.py:242-260) batched over leading dims; all float64.

    formulation="reference": inverse + logdets + matmul/trace exactly as written there.
    formulation="chol":      the algebraically identical Cholesky/triangular-solve form
                             (what the CUDA kernels evaluate) -- used as a cross-check.
    """
    T = m.shape[-1]
    if formulation == "reference":
        inv_p = torch.linalg.inv(Kp)
        p1 = torch.diagonal(inv_p @ Kq, dim1=-2, dim2=-1).sum(-1)
        p2 = _logdet(Kp) - _logdet(Kq)
        p3 = (m.unsqueeze(-2) @ (inv_p @ m.unsqueeze(-1))).squeeze(-1).squeeze(-1)
        return 0.5 * (p1 - T + p2 + p3)
    Lp = torch.linalg.cholesky(Kp)
    Lq = torch.linalg.cholesky(Kq)
    A = torch.linalg.solve_triangular(Lp, Lq, upper=False)
    a = torch.linalg.solve_triangular(Lp, m.unsqueeze(-1), upper=False).squeeze(-1)
    ld = 2.0 * (torch.log(torch.diagonal(Lp, dim1=-2, dim2=-1)).sum(-1)
                - torch.log(torch.diagonal(Lq, dim1=-2, dim2=-1)).sum(-1))
    return 0.5 * ((A * A).sum((-2, -1)) - T + ld + (a * a).sum(-1))


def kl_diag(m, logvar, K):
    """V2 gp_kl_div (VAE_GPprior_diag_cov.py:100-119) batched; float64 (log det via Cholesky)."""
    T = m.shape[-1]
    inv = torch.linalg.inv(K)
    v = torch.exp(logvar)
    p1 = (torch.diagonal(inv, dim1=-2, dim2=-1) * v).sum(-1)
    p2 = _logdet(K) - logvar.sum(-1)
    p3 = (m.unsqueeze(-2) @ (inv @ m.unsqueeze(-1))).squeeze(-1).squeeze(-1)
    return 0.5 * (p1 - T + p2 + p3)


def bidiag_dense(bd, bc):
    """Upper-bidiagonal B[..., T, T] with diagonal bd[..., T] and super-diagonal bc[..., :T-1]."""
    T = bd.shape[-1]
    Bm = torch.diag_embed(bd)
    if T > 1:
        Bm = Bm + torch.diag_embed(bc[..., : T - 1], offset=1)
    return Bm


def kl_bidiag(m, bd, bc, K):
    """V3 (extension): q = N(m, (B^T B)^-1), B upper bidiagonal."""
    T = m.shape[-1]
    Bm = bidiag_dense(bd, bc)
    Sigma = torch.linalg.inv(Bm.transpose(-1, -2) @ Bm)
    inv = torch.linalg.inv(K)
    p1 = torch.diagonal(inv @ Sigma, dim1=-2, dim2=-1).sum(-1)
    p2 = _logdet(K) + 2.0 * torch.log(bd).sum(-1)
    p3 = (m.unsqueeze(-2) @ (inv @ m.unsqueeze(-1))).squeeze(-1).squeeze(-1)
    return 0.5 * (p1 - T + p2 + p3)


def gp_prior_kl(mean, times, lengths, ell_q, ell_p, eps, *, aux=None, kernel=RBF, posterior="gp",
                noise=1e-3, S=1, formulation="reference", build_dtype=torch.float32):
    """Whole-batch oracle for the fused op (replaces prior_kernels/approx_kernels/gp_vae_sample/
    calc_gp_kl, Full_GP_VAE_dynamic_time.py:332-340).

    mean   [sum_T, D] f32   rows sequence-major then time (encoder output)
    times  [B, T_max] f32   zero padded on the right (DataHandler.py:151)
    lengths[B] int
    ell_q, ell_p [D] f32    posterior / prior lengthscales
    eps    [B, D, S, T_max] f32 standard-normal noise (explicit so results are reproducible)
    aux    posterior="diag":   logvar [sum_T, D]
           posterior="bidiag": [sum_T, D, 2]  (..,0)=diag b>0, (..,1)=super-diagonal c (last unused)
    Returns dict: z [S*sum_T, D] f64 (per sequence: S blocks of [T_b, D]), kl_sum (0-d f64),
    kl_pairs [B*D] f64 ordered p=b*D+d, logdet_p / logdet_q [B*D] f64.
    Tensors passed with requires_grad=True receive gradients through the returned values.
    """
    lengths_l = [int(x) for x in (lengths.tolist() if hasattr(lengths, "tolist") else lengths)]
    B = len(lengths_l)
    D = mean.shape[1]
    offs = np.concatenate([[0], np.cumsum(lengths_l)]).astype(np.int64)
    mean64 = mean.to(torch.float64)
    aux64 = None if aux is None else aux.to(torch.float64)
    kl_rows = [None] * B
    ldp_rows = [None] * B
    ldq_rows = [None] * B
    z_rows = [None] * B
    # group equal-length sequences so regular batches run as one batched LAPACK call
    groups = {}
    for b, T in enumerate(lengths_l):
        groups.setdefault(T, []).append(b)
    for T, bs in groups.items():
        if T == 0:
            for b in bs:
                kl_rows[b] = torch.zeros(D, dtype=torch.float64)
                ldp_rows[b] = torch.zeros(D, dtype=torch.float64)
                ldq_rows[b] = torch.zeros(D, dtype=torch.float64)
                z_rows[b] = torch.zeros(0, D, dtype=torch.float64)
            continue
        idx = torch.tensor(bs)
        t = times[idx][:, :T]                                   # [nb, T]
        rows = torch.stack([torch.arange(offs[b], offs[b] + T) for b in bs])  # [nb, T]
        m = mean64[rows].transpose(1, 2)                          # [nb, D, T]
        e = eps[idx][..., :T].to(torch.float64)                   # [nb, D, S, T]
        tt = t.unsqueeze(1).expand(len(bs), D, T)
        Kp = kernel_matrix(tt, ell_p.unsqueeze(0).expand(len(bs), D), kernel, noise, build_dtype)
        ldp = _logdet(Kp)
        if posterior == "gp":
            Kq = kernel_matrix(tt, ell_q.unsqueeze(0).expand(len(bs), D), kernel, noise, build_dtype)
            kl = kl_full_gp(m, Kq, Kp, formulation)
            Lq = torch.linalg.cholesky(Kq)
            ldq = 2.0 * torch.log(torch.diagonal(Lq, dim1=-2, dim2=-1)).sum(-1)
            zz = m.unsqueeze(2) + (Lq.unsqueeze(2) @ e.unsqueeze(-1)).squeeze(-1)   # [nb, D, S, T]
        elif posterior == "diag":
            lv = aux64[rows].transpose(1, 2)                      # [nb, D, T]
            kl = kl_diag(m, lv, Kp)
            ldq = lv.sum(-1)
            zz = m.unsqueeze(2) + torch.exp(0.5 * lv).unsqueeze(2) * e
        elif posterior == "bidiag":
            ab = aux64[rows]                                      # [nb, T, D, 2]
            bd = ab[..., 0].transpose(1, 2)
            bc = ab[..., 1].transpose(1, 2)
            kl = kl_bidiag(m, bd, bc, Kp)
            ldq = -2.0 * torch.log(bd).sum(-1)
            Bm = bidiag_dense(bd, bc)
            w = torch.linalg.solve_triangular(Bm.unsqueeze(2), e.unsqueeze(-1), upper=True).squeeze(-1)
            zz = m.unsqueeze(2) + w
        else:
            raise ValueError(posterior)
        for j, b in enumerate(bs):
            kl_rows[b] = kl[j]
            ldp_rows[b] = ldp[j]
            ldq_rows[b] = ldq[j]
            # [D, S, T] -> S blocks of [T, D]   (gp_vae_sample, Full_GP_VAE_dynamic_time.py:187-194)
            z_rows[b] = zz[j].permute(1, 2, 0).reshape(S * T, D)
    kl_pairs = torch.cat([k.reshape(-1) for k in kl_rows]) if B else torch.zeros(0, dtype=torch.float64)
    return {
        "z": torch.cat(z_rows, 0) if B else torch.zeros(0, D, dtype=torch.float64),
        "kl_sum": kl_pairs.sum(),
        "kl_pairs": kl_pairs,
        "logdet_p": torch.cat([k.reshape(-1) for k in ldp_rows]) if B else kl_pairs,
        "logdet_q": torch.cat([k.reshape(-1) for k in ldq_rows]) if B else kl_pairs,
    }


def gp_prior_kl_grads(mean, times, lengths, ell_q, ell_p, eps, g_z, g_kl_sum=1.0, g_kl_pairs=None, *,
                      aux=None, **kw):
    """Forward + autograd backward of  L = g_kl_sum*kl_sum + <g_kl_pairs, kl_pairs> + <g_z, z>.

    Returns (forward dict, dict of gradients: mean, ell_q, ell_p, aux) -- the stand-in for TF autodiff
    through the hot path (Full_GP_VAE_dynamic_time.py:361)."""
    mean = mean.detach().clone().requires_grad_(True)
    ell_q = ell_q.detach().clone().requires_grad_(True)
    ell_p = ell_p.detach().clone().requires_grad_(True)
    aux_r = None if aux is None else aux.detach().clone().requires_grad_(True)
    out = gp_prior_kl(mean, times, lengths, ell_q, ell_p, eps, aux=aux_r, **kw)
    loss = g_kl_sum * out["kl_sum"]
    if g_kl_pairs is not None:
        loss = loss + (g_kl_pairs.to(torch.float64) * out["kl_pairs"]).sum()
    if g_z is not None:
        loss = loss + (g_z.to(torch.float64) * out["z"]).sum()
    loss.backward()

    def _g(x):
        return None if x is None else (torch.zeros_like(x) if x.grad is None else x.grad.detach())
    grads = {"mean": _g(mean), "ell_q": _g(ell_q), "ell_p": _g(ell_p), "aux": _g(aux_r)}
    out = {k: v.detach() for k, v in out.items()}
    return out, grads


def bernoulli_recon(x, x_decode, lengths, S=1):
    """sum_recon_loss of the reference (Full_GP_VAE_dynamic_time.py:323-327 tiles x over the samples; :349 the
    Bernoulli NLL per row in float32; :350-356 float64, mean over samples, sum over time and batch).
    x [sum_T, F] f32, x_decode [S*sum_T, F] f32 laid out like z (per sequence S blocks of [T_b, F])."""
    lengths_l = [int(a) for a in lengths.tolist()]
    rows = []
    off = 0
    for T in lengths_l:
        rows += [x[off:off + T]] * S
        off += T
    x_tile = torch.cat(rows, 0) if rows else x[:0]
    rec = -(x_tile * torch.log(1e-10 + x_decode) + (1.0 - x_tile) * torch.log(1.0 - x_decode + 1e-10)).sum(1)
    return rec.to(torch.float64).sum() / S   # mean over the S samples of every (sequence, time) row, then the sum


def gp_recog_sample(mean, logvar, times, lengths, ell, eps, *, kernel=RBF, noise=1e-3, S=1, build_dtype=torch.float32):
    """GP-recognition sampler + standard KL (src/Models/GP_recog_VAE_prior.py): tf_kernel_approx builds K in float32, casts
    to float64, L = chol(K) + diag(sqrt(exp(logvar))) (:150-158), noise = L eps (:160-162); gp_vae_sample adds it to the
    mean (:170-191); standard_vae_kl (:65-70) per row of mean, cast to float64 and negated (:274-277).
    Returns dict z [S*sum_T, D] f64 (layout of gp_prior_kl), kl_rows [sum_T] f64, kl_sum."""
    lengths_l = [int(x) for x in (lengths.tolist() if hasattr(lengths, "tolist") else lengths)]
    B, D = len(lengths_l), mean.shape[1]
    offs = np.concatenate([[0], np.cumsum(lengths_l)]).astype(np.int64)
    mean64, lv64 = mean.to(torch.float64), logvar.to(torch.float64)
    var = torch.exp(lv64)
    kl_rows = -0.5 * (1.0 + torch.log(1e-10 + var) - mean64 ** 2 - var).sum(1)
    z_rows = []
    for b, T in enumerate(lengths_l):
        if T == 0:
            z_rows.append(torch.zeros(0, D, dtype=torch.float64))
            continue
        rows = slice(int(offs[b]), int(offs[b]) + T)
        t = times[b, :T].unsqueeze(0).expand(D, T)
        K = kernel_matrix(t, ell, kernel, noise, build_dtype)               # [D, T, T] float64 values of a float32 build
        L = torch.linalg.cholesky(K) + torch.diag_embed(torch.sqrt(var[rows].t()))
        e = eps[b][..., :T].to(torch.float64)                               # [D, S, T]
        zz = mean64[rows].t().unsqueeze(1) + (L.unsqueeze(1) @ e.unsqueeze(-1)).squeeze(-1)   # [D, S, T]
        z_rows.append(zz.permute(1, 2, 0).reshape(S * T, D))
    return {"z": torch.cat(z_rows, 0) if B else torch.zeros(0, D, dtype=torch.float64), "kl_rows": kl_rows,
            "kl_sum": kl_rows.sum()}


def gp_recog_grads(mean, logvar, times, lengths, ell, eps, g_z, g_kl_sum=1.0, g_kl_rows=None, **kw):
    """Forward + autograd backward of L = g_kl_sum*kl_sum + <g_kl_rows, kl_rows> + <g_z, z> (TF autodiff, :305)."""
    mean = mean.detach().clone().requires_grad_(True)
    logvar = logvar.detach().clone().requires_grad_(True)
    ell = ell.detach().clone().requires_grad_(True)
    out = gp_recog_sample(mean, logvar, times, lengths, ell, eps, **kw)
    loss = g_kl_sum * out["kl_sum"]
    if g_kl_rows is not None:
        loss = loss + (g_kl_rows.to(torch.float64) * out["kl_rows"]).sum()
    if g_z is not None:
        loss = loss + (g_z.to(torch.float64) * out["z"]).sum()
    loss.backward()
    grads = {"mean": mean.grad.detach(), "logvar": logvar.grad.detach(),
             "ell": torch.zeros_like(ell) if ell.grad is None else ell.grad.detach()}
    return {k: v.detach() for k, v in out.items()}, grads


def collate_batch(data, time_grid, index, max_time):
    """Ragged batch producer: numpy restatement of SyntheticDataHandler._prep_dataset (src/Models/DataHandler.py:129-156)
    for the sequences `index` (data_batch :111-127 slices consecutive ones): valid time points are those whose feature-0
    value is > -1 (:143), the kept values are packed [T_b, F] (:145), the time stamps zero padded to max_time (:149-151).
    data [N, F, T_full], time_grid [T_full].  Returns x [sum_T, F], times [B, max_time], lengths [B] int32."""
    xs, ts, ls = [], [], []
    for i in index:
        keep = np.where(data[i, 0, :] > -1)[0][:max_time]
        xs.append(data[i][:, keep].T)
        ts.append(np.pad(time_grid[keep], (0, max_time - len(keep)), "constant", constant_values=0))
        ls.append(len(keep))
    F = data.shape[1]
    x = np.concatenate(xs, 0) if xs else np.zeros((0, F), data.dtype)
    return x.astype(np.float32), np.stack(ts).astype(np.float32) if ts else np.zeros((0, max_time), np.float32), \
        np.asarray(ls, np.int32)



def posterior_impute(z_obs, t_obs, n_obs, t_full, eps=None, *, kernel=RBF, ell=1.0, noise=1e-3):
    """GP posterior conditioning of FullGP_and_GPdecoder_dynamic_time_analysis.py (sample_given_part_latent :40-56 looped
    as post_gp_sample :96-111), restated with numpy in the reference's own precisions: kernel entries float32 with the
    nugget wherever two time stamps coincide exactly (:8-22); L, Lk, mu float32 (:43-48); the posterior covariance
    K_ss + 1e-15 I - Lk^T Lk and its Cholesky float64 (:49-50).  z_obs [sum n_obs, D], t_obs [B, n_obs_max],
    n_obs [B], t_full [B, n_full], eps [B, D, n_full] or None (mean).  Returns (out [B*n_full, D] float64,
    failed [B] bool: covariance not positive definite -- the reference raises LinAlgError there)."""
    z_obs = np.asarray(z_obs, dtype=np.float32)
    t_obs, t_full = np.asarray(t_obs, dtype=np.float32), np.asarray(t_full, dtype=np.float32)
    n_obs = [int(x) for x in np.asarray(n_obs).tolist()]
    B, ns = t_full.shape
    D = z_obs.shape[1]
    offs = np.concatenate([[0], np.cumsum(n_obs)]).astype(np.int64)

    def kmat(a, b):
        d = a.astype(np.float64)[:, None] - b.astype(np.float64)[None, :]
        k = np.exp(-(d * d) / (2.0 * ell * ell)) if kernel == RBF else 1.0 / (1.0 + d * d / (ell * ell))
        nz = np.where(a[:, None] == b[None, :], noise, 0.0)
        return ((1.0 - nz) * k + nz).astype(np.float32)

    out = np.zeros((B * ns, D), dtype=np.float64)
    failed = np.zeros(B, dtype=bool)
    for b in range(B):
        td, tf = t_obs[b, : n_obs[b]], t_full[b]
        L = np.linalg.cholesky(kmat(td, td))                       # float32, :43-44
        Lk = np.linalg.solve(L, kmat(td, tf))                      # :46-47
        C = kmat(tf, tf) + 1e-15 * np.eye(ns) - np.dot(Lk.T, Lk)   # float64, :49-50
        try:
            Lc = np.linalg.cholesky(C)
        except np.linalg.LinAlgError:
            failed[b] = True
            Lc = np.full((ns, ns), np.nan)
        zb = z_obs[offs[b]: offs[b + 1]]                           # [n_obs, D]
        mu = np.dot(Lk.T, np.linalg.solve(L, zb))                  # [n_full, D], :48
        f = mu.astype(np.float64)
        if eps is not None:
            f = f + np.dot(Lc, np.asarray(eps[b], dtype=np.float64).T)   # :51
        out[b * ns: (b + 1) * ns] = f
    return out, failed


def gp_kl_div_numpy(m, Kq, Kp):
    """Single-pair numpy float64 transcription of gp_kl_div (Full_GP_VAE_dynamic_time.py:242-260),
    independent of torch -- second opinion for the golden tests."""
    m = np.asarray(m, np.float64).reshape(-1, 1)
    Kq = np.asarray(Kq, np.float64)
    Kp = np.asarray(Kp, np.float64)
    T = m.shape[0]
    inv_p = np.linalg.inv(Kp)
    ldq = 2.0 * np.log(np.diag(np.linalg.cholesky(Kq))).sum()
    ldp = 2.0 * np.log(np.diag(np.linalg.cholesky(Kp))).sum()
    p1 = np.trace(inv_p @ Kq)
    p3 = ((-m).T @ (inv_p @ (-m))).item()
    return 0.5 * (p1 - T + (ldp - ldq) + p3)


def synthetic_batch(B, D, T, S=1, *, ragged=False, seed=1234, posterior="gp", ell_p_value=1.0, grid=False):
    """Seeded synthetic inputs of SURVEY.md S8(d): irregular times cumsum(U(0.5,1.5)), mean/eps~N(0,1),
    l_p = ell_p_value, l_q = l_p*exp(N(0,0.1^2)); ragged: T_b ~ U{ceil(T/2)..T}.
    grid=True uses the reference's own time stamps 0,1,..,T-1 (DataHandler.py:42; cond(K) <= 68 at l=1)."""
    g = torch.Generator().manual_seed(seed)
    if ragged:
        lengths = torch.randint((T + 1) // 2, T + 1, (B,), generator=g, dtype=torch.int32)
        lengths[0] = T
    else:
        lengths = torch.full((B,), T, dtype=torch.int32)
    times = torch.cumsum(torch.rand(B, T, generator=g) + 0.5, dim=1).to(torch.float32)
    if grid:
        times = torch.arange(T, dtype=torch.float32).repeat(B, 1)
    for b in range(B):
        times[b, int(lengths[b]):] = 0.0
    total = int(lengths.sum())
    mean = torch.randn(total, D, generator=g, dtype=torch.float32)
    eps = torch.randn(B, D, S, T, generator=g, dtype=torch.float32)
    ell_p = torch.full((D,), float(ell_p_value), dtype=torch.float32)
    ell_q = (ell_p * torch.exp(0.1 * torch.randn(D, generator=g))).to(torch.float32)
    g_z = torch.randn(S * total, D, generator=g, dtype=torch.float32)
    aux = None
    if posterior == "diag":
        aux = (torch.rand(total, D, generator=g) * 2.5 - 2.0).to(torch.float32)
    elif posterior == "bidiag":
        bd = torch.nn.functional.softplus(torch.randn(total, D, generator=g)) + 0.1
        bc = 0.5 * torch.randn(total, D, generator=g)
        aux = torch.stack([bd, bc], -1).to(torch.float32)
    return dict(mean=mean, times=times, lengths=lengths, ell_q=ell_q, ell_p=ell_p, eps=eps, g_z=g_z, aux=aux)
