"""Coefficient matrices of Natural Inference and the per-step launch plan derived from them.

Formats follow the reference exactly:
  * npz with three arrays consumed BY POSITION -- past_xstart_coeff (K,K), past_epsilon_coeff
    (K,K) or (K,K+1), node_coeff (K+1,3) = (t, alpha, sigma)   (writer src/Utils.py:49, readers
    src/CIFAR10NaturalInference.py:273, src/ValidateNaturalInference.py:319);
  * SD3 csv: 28x28 table read with index_col=0 (src/SD3NaturalInference.py:196), row n-1 is
    normalised by its sum (:157-168) and mixed with sigma_{k+1}*noise (:209).
Host-side, O(K^2) scalars; no tensor arithmetic happens here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._lib import NI_MAX_TERMS


@dataclass
class CoeffTriple:
    """(A, B, node) with B normalised to K x (K+1) columns (column 0 = initial noise,
    column j>=1 = noise drawn at step j-1)."""
    A: np.ndarray
    B: np.ndarray
    node: np.ndarray
    name: str = ""

    def __post_init__(self):
        self.A = np.asarray(self.A, dtype=np.float64)
        self.B = np.asarray(self.B, dtype=np.float64)
        self.node = np.asarray(self.node, dtype=np.float64)
        K = self.A.shape[0]
        if self.A.shape != (K, K):
            raise ValueError(f"past_xstart_coeff must be square, got {self.A.shape}")
        if self.B.shape == (K, K):
            # weights/*.npz flavour.  Its reader (src/CIFAR10NaturalInference.py:303) uses column 0 only -- B[kk,0]*noise
            # with the single initial tensor -- so any other column is ignored here too (with a warning) instead of being
            # promoted to per-step fresh noise, which would silently run a different, stochastic sampler.
            if np.any(self.B[:, 1:] != 0):
                import warnings
                warnings.warn("K x K past_epsilon_coeff has non-zero columns >= 1; like the reference's CIFAR loop only column 0 "
                              "(the initial noise) is used", stacklevel=3)
            col0 = self.B[:, :1]
            self.B = np.concatenate([col0, np.zeros((K, K))], axis=1)
        if self.B.shape != (K, K + 1):
            raise ValueError(f"past_epsilon_coeff must be (K,K) or (K,K+1), got {self.B.shape}")
        if self.node.shape != (K + 1, 3):
            raise ValueError(f"node_coeff must be (K+1,3), got {self.node.shape}")
        if np.any(np.triu(self.A, 1) != 0):
            raise ValueError("past_xstart_coeff must be lower triangular (row k uses x0_0..x0_k)")
        if np.any(np.triu(self.B, 2) != 0):
            raise ValueError("past_epsilon_coeff row k may only use eps_0..eps_{k+1}")

    @property
    def K(self) -> int:
        return self.A.shape[0]

    @classmethod
    def from_npz(cls, path) -> "CoeffTriple":
        with np.load(path) as z:
            A, B, node = [z[k] for k in z.files]  # by position, like np.load(p).values()
        return cls(A, B, node, name=str(path))

    def save_npz(self, path):
        """Same array names and order as src/Utils.py:49."""
        np.savez(path, past_xstart_coeff=self.A, past_epsilon_coeff=self.B, node_coeff=self.node)

    def save_companion_csv(self, path):
        """The human-readable table `save_coeff_matrix` writes next to each npz (src/Utils.py:36-45): A rounded to 3
        decimals, columns = the K node times that produced each x0, index = the K times reached, plus a `sum` column.
        Times print as %03d on the discrete grid (mean t > 1) and %0.3f on the continuous one.  Written without pandas;
        byte-identical to the files under the reference's results/ (tests/test_host_logic.py)."""
        t = self.node[:, 0]
        names = [("%03d" % v) if t.mean() > 1 else ("%0.3f" % v) for v in t]
        cell = lambda v: repr(float(v))  # pandas' default float formatting is the shortest round-trip repr
        with open(path, "w", newline="") as f:
            f.write("," + ",".join(names[:-1]) + ",sum\n")
            for k in range(self.K):
                f.write(names[k + 1] + "," + ",".join(cell(v) for v in self.A[k].round(3)) + "," + cell(self.A[k].sum().round(3)) + "\n")

    @classmethod
    def from_sd3_table(cls, W, sigmas, name="") -> "CoeffTriple":
        """csv weight table + scheduler sigmas (length K+1, last = 0) -> common form:
        A[k,j] = (1-sigma_{k+1}) W[k,j]/sum_j W[k,j],  B[k,0] = sigma_{k+1}."""
        W = np.asarray(W, dtype=np.float64)
        sig = np.asarray(sigmas, dtype=np.float64)
        K = W.shape[0]
        if sig.shape != (K + 1,):
            raise ValueError("need K+1 sigmas")
        A = np.zeros((K, K))
        B = np.zeros((K, K + 1))
        for k in range(K):
            row = W[k, : k + 1]
            A[k, : k + 1] = (1.0 - sig[k + 1]) * row / row.sum()
            B[k, 0] = sig[k + 1]
        node = np.stack([sig, 1.0 - sig, sig], axis=1)
        return cls(A, B, node, name=name)

    @classmethod
    def from_sd3_csv(cls, path, sigmas=None) -> "CoeffTriple":
        W = load_weight_csv(path)
        if sigmas is None:
            sigmas = flow_match_sigmas(W.shape[0])
        return cls.from_sd3_table(W, sigmas, name=str(path))


def load_weight_csv(path) -> np.ndarray:
    """``pd.read_csv(path, index_col=0).to_numpy()`` without pandas."""
    with open(path, "r") as f:
        rows = [ln.strip().split(",") for ln in f if ln.strip()]
    return np.array([[float(v) for v in r[1:]] for r in rows[1:]], dtype=np.float64)


def save_weight_csv(W, sigmas, path):
    """Write an SD3 weight table in the layout of weights/sd3_step_28_weight*.csv (what `pd.read_csv(path, index_col=0)`
    at src/SD3NaturalInference.py:196 expects): header and index are sigma_1..sigma_K as %.2f, cells the shortest
    round-trip float repr.  `sigmas` has K+1 entries (sigma_0 = 1 first)."""
    W = np.asarray(W, dtype=np.float64)
    names = ["%.2f" % float(v) for v in list(sigmas)[1:]]
    if W.shape != (len(names), len(names)):
        raise ValueError(f"W is {W.shape}, expected ({len(names)}, {len(names)}) for {len(names) + 1} sigmas")
    with open(path, "w", newline="") as f:
        f.write("," + ",".join(names) + "\n")
        for k, row in enumerate(W):
            f.write(names[k] + "," + ",".join(repr(float(v)) for v in row) + "\n")


def flow_euler_weight_table(sigmas, cliplen: int = 0, rounded: bool = True) -> np.ndarray:
    """The table of plain flow-matching Euler in the csv's own units: W[k,j] = round(100 (sigma_j - sigma_{j+1}), 2) for
    j <= k.  With the FlowMatchEuler grid this IS weights/sd3_step_28_weight.csv (every cell); `euler_weighted_sum`
    (src/SD3NaturalInference.py:61-69) carries the same differences unrounded (rounded=False: W[k,j] = sigma_j - sigma_{j+1}).
    cliplen > 0 keeps only the last `cliplen` predictions of each row, the `seq_xstarts[-cliplen:]` window of
    `euler_weighted_sum(seq_xstarts, cliplen)` (:63, used at :118,:129,:131); 0 = the whole history, like the reference's slice."""
    sig = np.asarray(sigmas, dtype=np.float64)
    d = np.round(100.0 * (sig[:-1] - sig[1:]), 2) if rounded else sig[:-1] - sig[1:]
    W = np.tril(np.tile(d, (len(d), 1)))
    if cliplen > 0:
        W = W - np.tril(W, -int(cliplen))  # zero the columns j <= k - cliplen
    return W


def flow_match_sigmas(num_step=28, shift=3.0, num_train=1000) -> np.ndarray:
    """Sigma grid of diffusers' FlowMatchEulerDiscreteScheduler.set_timesteps (what
    src/SD3NaturalInference.py:188-190 reads from the pipeline), float32, trailing 0."""
    s_min = 1.0 / num_train
    sigma_min = shift * s_min / (1 + (shift - 1) * s_min)
    u = np.linspace(1.0, sigma_min, num_step, dtype=np.float32).astype(np.float64)
    return np.append(shift * u / (1 + (shift - 1) * u), 0.0).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# model I/O scaling: x0_k = a_k * x_k + b0_k * out0 + b1_k * out1
# ----------------------------------------------------------------------------------------------

def io_score_vp(node, beta_0=0.1, beta_1=20.0) -> List[Tuple[float, float, float]]:
    """CIFAR loop: out0 = raw net output h, score = -h/std(t) (deps/score_sde_pytorch/models/utils.py:150-159,
    sde_lib.py:141-145, fp32), x0 = (score*sigma^2 + x)/alpha (src/CIFAR10NaturalInference.py:229)."""
    out = []
    for k in range(node.shape[0] - 1):
        t = np.float32(node[k, 0])
        lmc = np.float32(-0.25) * t * t * np.float32(beta_1 - beta_0) - np.float32(0.5) * t * np.float32(beta_0)
        std = np.sqrt(np.float32(1.0) - np.exp(np.float32(2.0) * lmc, dtype=np.float32), dtype=np.float32)
        alpha, sigma = node[k, 1], node[k, 2]
        out.append((1.0 / alpha, -(sigma * sigma) / (alpha * float(std)), 0.0))
    return out


def io_eps_cfg(c1: Sequence[float], c2: Sequence[float], cfg_scale: Optional[float]) -> List[Tuple[float, float, float]]:
    """Validate loop: out0 = cond eps, out1 = uncond eps, fuse = uncond + s (cond - uncond)
    (src/ValidateNaturalInference.py:193), x0 = c1 z - c2 fuse (:355).  cfg_scale None -> single output."""
    out = []
    for a, c in zip(c1, c2):
        a, c = float(np.float32(a)), float(np.float32(c))
        if cfg_scale is None:
            out.append((a, -c, 0.0))
        else:
            out.append((a, -c * cfg_scale, -c * (1.0 - cfg_scale)))
    return out


def io_velocity_cfg(sigmas: Sequence[float], cfg_scale: Optional[float]) -> List[Tuple[float, float, float]]:
    """SD3 loop: out0 = v_text, out1 = v_null; x0 = x - sigma v, CFG applied on x0
    (src/SD3NaturalInference.py:215-217)."""
    out = []
    for s in list(sigmas)[:-1]:
        s = float(s)
        if cfg_scale is None:
            out.append((1.0, -s, 0.0))
        else:
            out.append((1.0, -s * cfg_scale, -s * (1.0 - cfg_scale)))
    return out


def ddim_x0_coeffs(num_step: int):
    """c1 = sqrt(1/abar), c2 = sqrt(1/abar - 1) on the sub-sampled grid, in SAMPLING order
    (src/ValidateNaturalInference.py:153-174, flipped at :328-329)."""
    idx = spaced_timesteps(1000, num_step)
    ab = np.cumprod(1.0 - np.linspace(0.0001, 0.02, 1000, dtype=np.float64))[idx]
    return np.sqrt(1.0 / ab)[::-1].copy(), np.sqrt(1.0 / ab - 1.0)[::-1].copy(), idx[::-1]


def spaced_timesteps(num_timesteps: int, count: int):
    """`space_timesteps(num_timesteps, str(count))` for one section (src/ValidateNaturalInference.py:57-78)."""
    stride = 1.0 if count <= 1 else (num_timesteps - 1) / (count - 1)
    cur, out = 0.0, []
    for _ in range(count):
        out.append(round(cur))
        cur += stride
    return sorted(set(out))


# ----------------------------------------------------------------------------------------------
# launch plan: which tensors row k reads, which it must keep, and where (ring slots)
# ----------------------------------------------------------------------------------------------

@dataclass
class StepPlanEntry:
    k: int
    c_x0: float                            # A[k,k]
    hist: List[Tuple[int, float]]          # (column j < k, A[k,j]) non-zero
    eps: List[Tuple[int, float]]           # (column j <= k, B[k,j]) non-zero: noise that already exists
    fresh: Optional[float]                 # B[k,k+1] if non-zero: noise drawn at this step
    keep_x0: bool                          # a later row reads x0_k
    keep_fresh: bool                       # a later row reads eps_{k+1}
    x0_slot: int = -1
    fresh_slot: int = -1
    c_xin: float = 0.0                     # Markov rows: coefficient of x_k standing for the whole history


@dataclass
class StepPlan:
    K: int
    steps: List[StepPlanEntry]
    n_x0_slots: int
    n_eps_slots: int                       # slots for eps_j, j >= 1 (eps_0 is separate)
    markov: bool = False                   # rows collapsed onto c_k * x_k (first-order samplers)
    x0_slot_of: List[int] = field(default_factory=list)
    eps_slot_of: List[int] = field(default_factory=list)   # index j (0 unused)
    eps0_last_use: int = -1

    def units(self, k: int, m_outputs: int, eps0_stored: bool = True, has_x_in: bool = True) -> int:
        """Tensor-sized HBM transfers of step k (BASELINE.md section 3 / SURVEY 8d):
        m + x_k + write x0_k (if kept) + history + stored noise + kept fresh noise + write x_{k+1}."""
        s = self.steps[k]
        u = m_outputs + (1 if has_x_in else 0) + (1 if s.keep_x0 else 0) + len(s.hist) + 1
        u += sum(1 for j, _ in s.eps if j != 0 or eps0_stored)
        u += 1 if (s.fresh is not None and s.keep_fresh) else 0
        return u

    def total_units(self, m_outputs: int, eps0_stored: bool = True) -> int:
        return sum(self.units(k, m_outputs, eps0_stored) for k in range(self.K))

    def launches(self, k: int, eps0_stored: bool = True) -> int:
        s = self.steps[k]
        n = len(s.hist) + sum(1 for j, _ in s.eps if j != 0 or eps0_stored)
        return max(1, -(-n // NI_MAX_TERMS))


def _alloc_slots(produce_step: Sequence[int], last_use: Sequence[int]):
    """Greedy ring allocation: a tensor produced at step p and last read at step l > p owns a slot
    during [p, l]; the slot is reusable from step l+1 on.  Returns (slot per tensor or -1, n_slots)."""
    order = sorted(range(len(produce_step)), key=lambda i: produce_step[i])
    free_at: List[int] = []  # per slot: first step at which it is free
    slot = [-1] * len(produce_step)
    for i in order:
        p, l = produce_step[i], last_use[i]
        if l <= p:
            continue
        for s_idx, f in enumerate(free_at):
            if f <= p:
                slot[i] = s_idx
                free_at[s_idx] = l + 1
                break
        else:
            slot[i] = len(free_at)
            free_at.append(l + 1)
    return slot, len(free_at)


def markov_ratios(triple: CoeffTriple, tol: float = 1e-12):
    """First-order structure of the x0 part.  If every row satisfies A[k,:k] = c_k * A[k-1,:k] (DDPM, DDIM,
    Euler-Maruyama, probability-flow Euler, flow-matching Euler, the default SD3 table) return (c, R) where
        sum_{j<k} A[k,j] x0_j + sum_{j<=k} B[k,j] eps_j  =  c_k * x_k + sum_{j<=k} R[k,j] eps_j,
        R[k,j] = B[k,j] - c_k * B[k-1,j]     (row -1 := the unit vector on eps_0, because x_0 = eps_0)
    so a step needs no x0 history at all, and noise only where R is non-zero (nowhere for exact first-order
    samplers; eps_0 for the 2-decimal SD3 table).  Returns None when the x0 part is not first-order."""
    A, B, K = triple.A, triple.B, triple.K
    cs, R = [], np.zeros((K, K + 1))
    prev_b = np.zeros(K + 1)
    prev_b[0] = 1.0
    for k in range(K):
        prev_a, cur_a = (A[k - 1, :k], A[k, :k]) if k > 0 else (np.zeros(0), np.zeros(0))
        nz = np.abs(prev_a) > 0
        if nz.any():
            c = float(np.median(cur_a[nz] / prev_a[nz]))
        else:  # no x0 column to pin the ratio: take it from the noise part
            nzb = np.abs(prev_b[: k + 1]) > 0
            c = float(np.median(B[k, : k + 1][nzb] / prev_b[: k + 1][nzb])) if nzb.any() else 0.0
        if k > 0 and np.abs(cur_a - c * prev_a).max(initial=0.0) > tol * max(1.0, np.abs(cur_a).max(initial=0.0)):
            return None
        r = B[k, : k + 1] - c * prev_b[: k + 1]
        r[np.abs(r) <= tol * max(1.0, np.abs(B[k]).max())] = 0.0
        R[k, : k + 1] = r
        cs.append(c)
        prev_b = B[k].copy()
    return cs, R


def build_plan(triple: CoeffTriple, keep_all_x0: bool = False, markov: bool = False) -> StepPlan:
    A, B, K = triple.A, triple.B, triple.K
    if markov:
        mr = markov_ratios(triple)
        if mr is None:
            raise ValueError("the x0 part of the matrix is not first-order (Markov); build the plan with markov=False")
        cs, R = mr
        nzR = R != 0
        eps_last = [max([k for k in range(K) if nzR[k, j]], default=-1) for j in range(K + 1)]
        eps_slot_j, n_eps = _alloc_slots([j - 1 for j in range(1, K + 1)], eps_last[1:])
        eps_slot = [-1] + eps_slot_j
        x0_slot = list(range(K)) if keep_all_x0 else [-1] * K
        steps = []
        for k in range(K):
            fresh = float(B[k, k + 1]) if B[k, k + 1] != 0 else None
            keep_fresh = eps_last[k + 1] > k
            if keep_fresh and fresh is None:
                fresh = 0.0
            steps.append(StepPlanEntry(k=k, c_x0=float(A[k, k]), hist=[], eps=[(j, float(R[k, j])) for j in range(k + 1) if nzR[k, j]],
                                       fresh=fresh, keep_x0=keep_all_x0, keep_fresh=keep_fresh, x0_slot=x0_slot[k],
                                       fresh_slot=eps_slot[k + 1], c_xin=cs[k]))
        return StepPlan(K=K, steps=steps, n_x0_slots=K if keep_all_x0 else 0, n_eps_slots=n_eps, x0_slot_of=x0_slot,
                        eps_slot_of=eps_slot, eps0_last_use=eps_last[0], markov=True)
    nzA = A != 0
    nzB = B != 0
    # last row that reads column j
    x0_last = [max([k for k in range(K) if nzA[k, j]], default=-1) for j in range(K)]
    eps_last = [max([k for k in range(K) if nzB[k, j]], default=-1) for j in range(K + 1)]
    if keep_all_x0:
        x0_last = [K] * K
    x0_slot, n_x0 = _alloc_slots(list(range(K)), x0_last)
    # eps_j (j>=1) is produced at step j-1
    eps_slot_j, n_eps = _alloc_slots([j - 1 for j in range(1, K + 1)], eps_last[1:])
    eps_slot = [-1] + eps_slot_j
    steps = []
    for k in range(K):
        hist = [(j, float(A[k, j])) for j in range(k) if nzA[k, j]]
        eps = [(j, float(B[k, j])) for j in range(k + 1) if nzB[k, j]]
        fresh = float(B[k, k + 1]) if nzB[k, k + 1] else None
        keep_fresh = eps_last[k + 1] > k
        if keep_fresh and fresh is None:
            fresh = 0.0  # a later row uses eps_{k+1} although this row's coefficient is zero: still drawn here
        steps.append(StepPlanEntry(k=k, c_x0=float(A[k, k]), hist=hist, eps=eps, fresh=fresh,
                                   keep_x0=x0_last[k] > k, keep_fresh=keep_fresh,
                                   x0_slot=x0_slot[k], fresh_slot=eps_slot[k + 1]))
    return StepPlan(K=K, steps=steps, n_x0_slots=n_x0, n_eps_slots=n_eps, x0_slot_of=x0_slot,
                    eps_slot_of=eps_slot, eps0_last_use=eps_last[0])
