"""Host-side mirror of the hot path: plans, device runs and the reference-facing host entry points.

``FeaturePlan`` wraps the opaque ``CvPlan`` of the C-ABI (include/caviar_b200.h).  PyTorch is used only for
device memory and streams; every computation happens in the sm_100a kernels behind the C-ABI.
"""
from __future__ import annotations

import ctypes as C
import itertools
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (BOX_NONE, BOX_ORTHO, BOX_TRICLINIC, CvAggregates, CvFrameOutputs, CvHostOutputs, CvPlanDesc,
                   CvPlanInfo, MODE_ALLPAIRS, MODE_DIRECT, MODE_STAGED, MODE_ZEROCOPY, ORDER_CALLER, ORDER_PLAN, PLAN_FORCE_DIRECT,
                   PLAN_FORCE_STAGED, check)

_BOX_KINDS = {None: BOX_NONE, "none": BOX_NONE, "ortho": BOX_ORTHO, "orthorhombic": BOX_ORTHO,
              "triclinic": BOX_TRICLINIC}
_FORCE = {None: 0, "auto": 0, "direct": PLAN_FORCE_DIRECT, "staged": PLAN_FORCE_STAGED}

# Cut-off defaults.  8.0 is the reference's cluster_cut_A (CAVIAR-1.0.py:62); the reference has no angle or
# H-bond criterion, 3.0 A / 135 deg are cpptraj's `hbond` defaults and are an extension of ours.
DEFAULT_PAIR_CUTOFF = 8.0
DEFAULT_HBOND_DIST_CUTOFF = 3.0
DEFAULT_HBOND_ANGLE_CUTOFF = 135.0


def _index_array(rows, width: int, n_atoms: int, what: str) -> np.ndarray:
    """(K, width) int32, C-contiguous; negative indices wrap like numpy fancy indexing (CAVIAR-1.0.py:118)."""
    if rows is None:
        return np.zeros((0, width), dtype=np.int32)
    arr = None
    if type(rows) is list and rows and type(rows[0]) is tuple:
        # the reference's own input: a long Python list of int tuples (CAVIAR-1.0.py:331).  np.asarray() spends
        # 18 ms on 43 660 tuples; a flat iterator over verified plain ints takes half of that.
        try:
            if set(map(len, rows)) == {width} and set(map(type, itertools.chain.from_iterable(rows))) == {int}:
                arr = np.fromiter(itertools.chain.from_iterable(rows), dtype=np.int64,
                                  count=width * len(rows)).reshape(-1, width)
        except (TypeError, OverflowError):
            arr = None
    if arr is None:
        arr = np.asarray(rows)
    if arr.size == 0:
        return np.zeros((0, width), dtype=np.int32)
    if arr.ndim != 2 or arr.shape[1] != width:
        raise ValueError(f"{what} must have shape (K, {width}), got {arr.shape}")
    if not np.issubdtype(arr.dtype, np.integer):
        raise IndexError(f"{what} must be integers")
    arr = arr.astype(np.int64, copy=True)
    bad = (arr < -n_atoms) | (arr >= n_atoms)
    if bad.any():
        k = int(arr[bad][0])
        raise IndexError(f"index {k} is out of bounds for axis 1 with size {n_atoms}")
    arr[arr < 0] += n_atoms
    return np.ascontiguousarray(arr, dtype=np.int32)


def normalize_box(box, n_frames: int, kind: int) -> Optional[np.ndarray]:
    """None | (3,) | (T,3) | (3,3) | (T,3,3) -> float32 (T,3) [ortho] or (T,9) [triclinic] as the C-ABI wants."""
    if kind == BOX_NONE:
        return None
    if box is None:
        raise ValueError("plan has a periodic box but no box was given")
    b = np.asarray(box, dtype=np.float64)
    if kind == BOX_ORTHO:
        if b.shape == (3,):
            b = np.broadcast_to(b, (n_frames, 3))
        if b.shape != (n_frames, 3):
            raise ValueError(f"orthorhombic box must be (3,) or ({n_frames}, 3), got {b.shape}")
        return np.ascontiguousarray(b, dtype=np.float32)
    if b.shape == (3, 3):
        b = np.broadcast_to(b, (n_frames, 3, 3))
    if b.shape == (n_frames, 9):
        b = b.reshape(n_frames, 3, 3)
    if b.shape != (n_frames, 3, 3):
        raise ValueError(f"triclinic box must be (3,3) or ({n_frames}, 3, 3), got {b.shape}")
    return np.ascontiguousarray(b.reshape(n_frames, 9), dtype=np.float32)


@dataclass
class FrameOutputs:
    dist: object = None      # (T, P) float32
    angle: object = None     # (T, A) float32, degrees
    flags: object = None     # (T, W) packed bits (int32 storage on device, uint32 on host)
    score: object = None     # (T,) int32


@dataclass
class Aggregates:
    occupancy: object        # (P + A,) int64 (device) / uint64 (host)
    sum: object              # (P + A,) float64
    sumsq: object            # (P + A,) float64

    def mean(self, n_frames: int):
        return self.sum / float(n_frames)

    def var(self, n_frames: int, ddof: int = 1):
        """Per-feature variance from the running sums (CAVIAR-1.0.py:539-545 uses ddof=1)."""
        m = self.sum / float(n_frames)
        return (self.sumsq - n_frames * m * m) / float(n_frames - ddof)


class FeaturePlan:
    """Pair / triplet lists + cut-offs compiled into a persistent-kernel work plan on one GPU."""

    def __init__(self, n_atoms: int, pairs=None, triplets=None, box: Optional[str] = None,
                 pair_cutoff: float = DEFAULT_PAIR_CUTOFF, hbond_dist_cutoff: float = DEFAULT_HBOND_DIST_CUTOFF,
                 hbond_angle_cutoff: float = DEFAULT_HBOND_ANGLE_CUTOFF, force: Optional[str] = None,
                 device: Optional[int] = None, n_ctas: int = 0, describe_only_sms: Optional[int] = None):
        self.n_atoms = int(n_atoms)
        self.pairs = _index_array(pairs, 2, self.n_atoms, "pairs")
        self.triplets = _index_array(triplets, 3, self.n_atoms, "triplets")
        self.box_kind = _BOX_KINDS[box]
        self.pair_cutoff = float(pair_cutoff)
        self.hbond_dist_cutoff = float(hbond_dist_cutoff)
        self.hbond_angle_cutoff = float(hbond_angle_cutoff)
        self._handle = C.c_void_p()
        desc = CvPlanDesc()
        desc.struct_size = C.sizeof(CvPlanDesc)
        desc.n_atoms = self.n_atoms
        desc.n_pairs = len(self.pairs)
        desc.pairs = self.pairs.ctypes.data if len(self.pairs) else None
        desc.n_triplets = len(self.triplets)
        desc.triplets = self.triplets.ctypes.data if len(self.triplets) else None
        desc.box_kind = self.box_kind
        desc.flags = _FORCE[force]
        desc.pair_cutoff = self.pair_cutoff
        desc.hbond_dist_cutoff = self.hbond_dist_cutoff
        desc.hbond_angle_cutoff = self.hbond_angle_cutoff
        desc.device = -1 if device is None else int(device)
        desc.n_ctas = int(n_ctas)
        info = CvPlanInfo()
        self._desc, self._describe_sms = desc, describe_only_sms
        if describe_only_sms is not None:      # host-only dry run (no CUDA): used by CPU tests
            check(_lib.lib.caviar_plan_describe(C.byref(desc), int(describe_only_sms), C.byref(info)))
        else:
            check(_lib.lib.caviar_plan_create(C.byref(desc), C.byref(self._handle)))
            check(_lib.lib.caviar_plan_info(self._handle, C.byref(info)))
        self.info = info.as_dict()
        # the layout accessors re-run the (deterministic) host planner; the plan order does not depend on the SM count
        self._sms = int(describe_only_sms) if describe_only_sms is not None else 148
        self.P, self.A = len(self.pairs), len(self.triplets)
        self.F = self.P + self.A
        self.W = int(info.n_flag_words)
        self.mode = int(info.mode)
        self.device = device

    def allpairs_check(self) -> dict:
        """Host-side verification of the all-pairs geometry (``mode_name == "allpairs"``): the C side runs the kernel's
        own per-thread set-up for every tile, warp and lane and checks that every pair is evaluated exactly once from
        the right atoms.  Raises on a violation; returns the geometry statistics (``n_tiles`` = 0: flavour not used)."""
        stats = np.zeros(8, dtype=np.int64)
        got = _lib.lib.caviar_plan_allpairs_check(C.byref(self._desc), int(self._sms), stats.ctypes.data)
        if got < 0:
            check(int(got))
        keys = ("windows", "fast_windows", "max_tile_atoms", "frames_per_stage", "stages", "smem_bytes", "mean_tile_atoms", "n_ctas")
        return {"n_tiles": int(got), **{k: int(v) for k, v in zip(keys, stats)}}

    def chunks(self, n_frames: int) -> dict:
        """How ``run`` would cut ``n_frames`` frames into the frame chunks of its work queue (host-side, no CUDA)."""
        out = np.zeros(4, dtype=np.int32)
        check(_lib.lib.caviar_plan_chunks(C.byref(self._desc), self._sms, int(n_frames), out.ctypes.data))
        return {"chunk_frames": int(out[0]), "tail_frames": int(out[1]), "n_big_chunks": int(out[2]), "n_chunks": int(out[3])}

    def layout(self) -> dict:
        """The planner's tables (host-side, no CUDA): ``tiles`` (n_tiles, 8) = kind, f_begin, n_feat, idx_off, grp_off,
        grp_floats, 0, 0; ``idx`` (3, n_tiles*896) float offsets inside a tile's block; ``src_off`` (S,)."""
        n_tiles, S = int(self.info["n_tiles"]), int(self.info["staged_floats_per_frame"])
        tiles = np.zeros((n_tiles, 8), dtype=np.int32)
        idx = np.zeros((3, n_tiles * 896), dtype=np.int32)
        src = np.zeros(S, dtype=np.int32)
        check(_lib.lib.caviar_plan_layout(C.byref(self._desc), int(self._describe_sms or 148), tiles.ctypes.data,
                                          idx.ctypes.data, src.ctypes.data if S else None))
        return {"tiles": tiles, "idx": idx, "src_off": src}

    @property
    def atom_order(self) -> Optional[np.ndarray]:
        """The PLAN ORDER: (S/3,) int32, slot k of a plan-order frame holds atom ``atom_order[k]`` of the caller's
        numbering (per-tile blocks; an atom referenced by several tiles is repeated, block tails are padded).
        ``frames[:, plan.atom_order, :]`` is the hand-over format of staged plans -- what ``atom_slice(...).xyz``
        (CAVIAR-1.0.py:335) produces when it is given this list.  ``None`` for plans that are not staged (those read
        the caller's frames as they are)."""
        if self.mode != MODE_STAGED:
            return None
        if getattr(self, "_atom_order", None) is None:
            n = int(self.info["staged_floats_per_frame"]) // 3
            order = np.zeros(n, dtype=np.int32)
            got = _lib.lib.caviar_plan_atom_order(C.byref(self._desc), int(self._sms), order.ctypes.data, n)
            if got < 0:
                check(int(got))
            assert got == n
            self._atom_order = order
        return self._atom_order

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            _lib.lib.caviar_plan_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- device path ------------------------------------------------------------------------------
    @staticmethod
    def _stream_ptr():
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def prepare_box(self, box):
        """Raw per-frame boxes (torch CUDA float32 (T,3) | (T,9) | (T,3,3)) -> box records (T, boxrec_floats_per_frame)."""
        import torch
        if self.box_kind == BOX_NONE:
            return None
        width = 3 if self.box_kind == BOX_ORTHO else 9
        b = box.reshape(box.shape[0], -1).to(torch.float32).contiguous()
        if b.shape[1] != width:
            raise ValueError(f"box must have {width} floats per frame, got {b.shape[1]}")
        rec = torch.empty((b.shape[0], int(self.info["boxrec_floats_per_frame"])), dtype=torch.float32, device=b.device)
        check(_lib.lib.caviar_prepare_box(self._handle, b.data_ptr(), b.shape[0], rec.data_ptr(), self._stream_ptr()))
        return rec

    def stage(self, xyz, boxrec=None, out=None):
        """K0: (T, N, 3) CUDA float32 frames in the caller's atom order -> plan-order (T, S) layout, i.e.
        ``xyz[:, plan.atom_order, :]`` on the device (bit-for-bit copies).  Identity for non-staged plans.
        ``boxrec`` is accepted for source compatibility and ignored: staging is pure data movement."""
        import torch
        if self.mode == MODE_ALLPAIRS:
            # the all-pairs kernel streams ranges of the caller's frames: frames a multiple of 16 bytes apart and
            # 3 * ceil4(n_atoms) floats long.  Anything else (a selection size that is not a multiple of 4) is copied
            # once into padded frames; the view handed back has the caller's shape.
            need = 3 * ((self.n_atoms + 3) // 4 * 4)
            if xyz.dim() == 3 and xyz.shape[0] > 1 and xyz.stride(0) % 4 == 0 and xyz.stride(0) >= need and xyz.data_ptr() % 16 == 0:
                return xyz
            if xyz.dim() == 3 and xyz.shape[0] <= 1 and need == 3 * self.n_atoms and xyz.data_ptr() % 16 == 0:
                return xyz
            padded = torch.zeros((xyz.shape[0], need // 3, 3), dtype=torch.float32, device=xyz.device)
            padded[:, :self.n_atoms].copy_(xyz)
            return padded[:, :self.n_atoms]
        if self.mode != MODE_STAGED:
            return xyz
        if xyz.dtype != torch.float32 or not xyz.is_cuda:
            raise TypeError("stage() wants a CUDA float32 tensor")
        if xyz.dim() != 3 or xyz.shape[1] != self.n_atoms or xyz.shape[2] != 3 or xyz.stride(2) != 1 or xyz.stride(1) != 3:
            raise ValueError(f"frames must be (T, {self.n_atoms}, 3) with contiguous atoms")
        T = xyz.shape[0]
        S = int(self.info["staged_floats_per_frame"])
        if out is None:
            out = torch.empty((T, S), dtype=torch.float32, device=xyz.device)
        check(_lib.lib.caviar_stage_frames(self._handle, xyz.data_ptr(), xyz.stride(0) if T > 1 else 3 * self.n_atoms,
                                           T, out.data_ptr(), self._stream_ptr()))
        self.last_launches = int(_lib.lib.caviar_last_launch_count())
        return out

    def new_aggregates(self, device="cuda", moments: bool = True) -> Aggregates:
        """Zeroed running aggregates; ``moments=False`` keeps the occupancy only (contact-map runs)."""
        import torch
        if not moments:
            return Aggregates(torch.zeros(self.F, dtype=torch.int64, device=device), None, None)
        return Aggregates(torch.zeros(self.F, dtype=torch.int64, device=device),
                          torch.zeros(self.F, dtype=torch.float64, device=device),
                          torch.zeros(self.F, dtype=torch.float64, device=device))

    def new_packed_aggregates(self, device="cuda", moments: bool = True):
        """Aggregates in ONE contiguous float64 buffer ``[occupancy | sum | sumsq]`` (3F, or F without the moments):
        returns ``(buffer, Aggregates)`` whose fields are views of it.  Frame counts are exact in float64 below 2^53, so
        the multi-GPU exchange step of the path is a single ``all_reduce(buffer)`` -- no second collective for an
        integer tensor, nothing to pack or unpack -- and the reduced occupancy is still an exact integer."""
        import torch
        buf = torch.zeros((3 if moments else 1) * self.F, dtype=torch.float64, device=device)
        F = self.F
        return buf, Aggregates(buf[:F], buf[F:2 * F] if moments else None, buf[2 * F:] if moments else None)

    def new_workspace(self, device="cuda"):
        import torch
        return torch.empty(max(int(self.info["workspace_bytes"]), 8), dtype=torch.uint8, device=device)

    def alloc_outputs(self, n_frames: int, want: Sequence[str] = ("dist", "angle", "flags", "score"), device="cuda"):
        import torch
        o = FrameOutputs()
        if "dist" in want and self.P:
            o.dist = torch.empty((n_frames, self.P), dtype=torch.float32, device=device)
        if "angle" in want and self.A:
            o.angle = torch.empty((n_frames, self.A), dtype=torch.float32, device=device)
        if "flags" in want:
            o.flags = torch.empty((n_frames, self.W), dtype=torch.int32, device=device)
        if "score" in want:
            o.score = torch.empty((n_frames,), dtype=torch.int32, device=device)
        return o

    def run(self, coords, boxrec=None, out: Optional[FrameOutputs] = None, aggregates: Optional[Aggregates] = None,
            workspace=None, n_frames: Optional[int] = None) -> FrameOutputs:
        """The fused kernel over frames resident on the device (staged layout for staged plans).

        Asynchronous on torch's current stream.  ``aggregates`` are accumulated into.
        """
        import torch
        if coords.dtype != torch.float32 or not coords.is_cuda:
            raise TypeError("run() wants a CUDA float32 tensor")
        T = coords.shape[0] if n_frames is None else int(n_frames)
        if self.mode == MODE_STAGED:
            S = int(self.info["staged_floats_per_frame"])
            ok = (coords.dim() == 2 and coords.shape[1] == S and coords.stride(1) == 1) or \
                 (coords.dim() == 3 and coords.shape[1] * 3 == S and coords.shape[2] == 3 and coords.stride(2) == 1
                  and coords.stride(1) == 3)
            if not ok:
                raise ValueError(f"staged plan: pass frames in plan order -- (T, {S // 3}, 3) = frames[:, plan.atom_order, :], "
                                 "or the (T, S) tensor returned by stage()")
            stride = coords.stride(0) if coords.shape[0] > 1 else S
        else:
            if coords.dim() != 3 or coords.shape[1] != self.n_atoms or coords.shape[2] != 3 \
                    or coords.stride(2) != 1 or coords.stride(1) != 3:
                raise ValueError(f"frames must be (T, {self.n_atoms}, 3) with contiguous atoms")
            stride = coords.stride(0) if (coords.shape[0] > 1 or coords.stride(0) >= 3 * self.n_atoms) else 3 * self.n_atoms
        if out is None:
            out = self.alloc_outputs(T, device=coords.device)
        fo = CvFrameOutputs(*(None if t is None else t.data_ptr() for t in (out.dist, out.angle, out.flags, out.score)))
        ag_ptr = None
        if workspace is None:      # work queue + partial aggregates; one cached buffer per plan (single-stream use)
            if getattr(self, "_ws", None) is None or self._ws.device != coords.device:
                self._ws = self.new_workspace(coords.device)
            workspace = self._ws
        ws_ptr = workspace.data_ptr()
        if aggregates is not None:      # occupancy alone (sum = sumsq = None), or all three
            occ = aggregates.occupancy
            if occ.dtype not in (torch.int64, torch.float64) or not occ.is_contiguous() or occ.shape != (self.F,):
                raise ValueError(f"occupancy must be a contiguous ({self.F},) int64 or float64 CUDA tensor")
            f64 = occ.dtype == torch.float64        # counts as float64 (exact < 2^53): see new_packed_aggregates
            ag = CvAggregates(None if f64 else occ.data_ptr(),
                              None if aggregates.sum is None else aggregates.sum.data_ptr(),
                              None if aggregates.sumsq is None else aggregates.sumsq.data_ptr(),
                              occ.data_ptr() if f64 else None)
            ag_ptr = C.byref(ag)
        check(_lib.lib.caviar_run(self._handle, coords.data_ptr(), stride,
                                  None if boxrec is None else boxrec.data_ptr(), T,
                                  C.byref(fo), ag_ptr, ws_ptr, self._stream_ptr()))
        self.last_launches = int(_lib.lib.caviar_last_launch_count())
        return out

    def capture(self, coords, boxrec=None, out: Optional[FrameOutputs] = None, aggregates: Optional[Aggregates] = None,
                workspace=None, n_frames: Optional[int] = None):
        """Record one pass of ``run`` over these buffers -- its two memsets, the fused kernel and the finalize kernel --
        as a CUDA graph.  ``graph.replay()`` then repeats the pass for one launch's worth of host time: short passes
        (BASELINE config 1: ~20 us of device time) are otherwise bound by the host enqueueing four stream operations.
        Nothing is executed on the caller's buffers except ``out`` (written once by the warm-up, as a replay would);
        ``aggregates`` are accumulated into by every replay, not by this call.  Returns ``(graph, out)``."""
        import torch
        T = coords.shape[0] if n_frames is None else int(n_frames)
        if out is None:
            out = self.alloc_outputs(T, device=coords.device)
        if workspace is None:
            if getattr(self, "_ws", None) is None or self._ws.device != coords.device:
                self._ws = self.new_workspace(coords.device)
            workspace = self._ws
        scratch = None
        if aggregates is not None:      # same kernel flavour as the recorded pass, on throw-away totals
            scratch = Aggregates(torch.zeros_like(aggregates.occupancy),
                                 None if aggregates.sum is None else torch.zeros_like(aggregates.sum),
                                 None if aggregates.sumsq is None else torch.zeros_like(aggregates.sumsq))
        side = torch.cuda.Stream(device=coords.device)
        side.wait_stream(torch.cuda.current_stream(coords.device))
        with torch.cuda.stream(side):   # one-time set-up (shared-memory opt-in of the flavour) must not fall into the capture
            self.run(coords, boxrec, out=out, aggregates=scratch, workspace=workspace, n_frames=T)
        torch.cuda.current_stream(coords.device).wait_stream(side)
        torch.cuda.synchronize(coords.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.run(coords, boxrec, out=out, aggregates=aggregates, workspace=workspace, n_frames=T)
        return graph, out

    def residue_components(self, sum_a, n_frames: int, cut: float, sum_b=None, n_frames_b: int = 0):
        """Component label of every residue (= atom of this plan) on the device: ``build_residue_components``
        (CAVIAR-1.0.py:183-199) over the per-pair time means ``sum / T`` -- or, with ``sum_b``, the pooled mean of two
        ensembles (:496) -- thresholded strictly below ``cut`` in double (:186).  ``sum_a`` / ``sum_b``: the fp64 ``sum``
        aggregates of ``run`` (CUDA tensors).  Returns an int32 CUDA tensor of n_atoms labels, numbered like the
        reference's (in the order of each component's lowest residue)."""
        import torch
        for t in (sum_a, sum_b):
            if t is not None and (not t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous() or t.shape[0] < self.P):
                raise ValueError("sums must be contiguous float64 CUDA tensors with at least P entries")
        labels = torch.empty(self.n_atoms, dtype=torch.int32, device=sum_a.device)
        check(_lib.lib.caviar_residue_components(self._handle, sum_a.data_ptr(), int(n_frames),
                                                 None if sum_b is None else sum_b.data_ptr(), int(n_frames_b), float(cut),
                                                 labels.data_ptr(), self._stream_ptr()))
        return labels

    def weighted_scores(self, flags, weights):
        """Per-frame weighted interaction score from the packed flags of ``run``: ``sum_f w[f] * flag[t, f]``.

        ``weights``: (P + A,) array-like, pairs first then triplets -- e.g. the ``abs_loading`` column CAVIAR writes
        next to each pair (CAVIAR-1.0.py:141-147).  They are rounded to 32.32 fixed point, summed exactly in int64
        on the device (bit-reproducible, any summation order) and returned as a float64 CUDA tensor of shape (T,).
        """
        import torch
        wq = quantize_weights(weights, self.F)
        if flags.dtype not in (torch.int32, torch.uint32) or not flags.is_cuda or flags.dim() != 2 or flags.shape[1] != self.W \
                or not flags.is_contiguous():
            raise ValueError(f"flags must be the contiguous (T, {self.W}) int32 CUDA tensor written by run()")
        wq_dev = torch.from_numpy(wq).to(flags.device)
        out = torch.empty(flags.shape[0], dtype=torch.int64, device=flags.device)
        check(_lib.lib.caviar_weighted_scores(self._handle, flags.data_ptr(), flags.shape[0], wq_dev.data_ptr(),
                                              out.data_ptr(), self._stream_ptr()))
        return out.to(torch.float64) / 4294967296.0

    # -- reference-facing host path ---------------------------------------------------------------
    def features_host(self, xyz: np.ndarray, box=None, want: Sequence[str] = ("dist", "angle", "flags", "score", "agg"),
                      frames_per_block: int = 0, out: Optional[dict] = None, weights=None, plan_order: bool = False) -> dict:
        """HOST arrays in, HOST arrays out (H2D, kernels, D2H inside).  ``xyz`` is (T, N, 3) float32 in the caller's atom
        order, or -- ``plan_order=True``, staged plans -- (T, len(plan.atom_order), 3) already sliced in plan order
        (``frames[:, plan.atom_order, :]``), which skips the device-side permutation.

        ``weights`` (one per feature, pairs first) adds ``res["score_weighted"]``: float64 (T,), the per-frame sum of
        the weights of the set flags (exact 32.32 fixed-point sum on the device, see ``weighted_scores``)."""
        xyz = np.asarray(xyz)
        plan_order = bool(plan_order) and self.mode == MODE_STAGED
        n_in = int(self.info["staged_floats_per_frame"]) // 3 if plan_order else self.n_atoms
        if xyz.ndim != 3 or xyz.shape[1] != n_in or xyz.shape[2] != 3:
            raise ValueError(f"coordinates must be (T, {n_in}, 3), got {xyz.shape}")
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        T = xyz.shape[0]
        b = normalize_box(box, T, self.box_kind)
        res = {} if out is None else out
        if "dist" in want and self.P and "dist" not in res:
            res["dist"] = np.empty((T, self.P), dtype=np.float32)
        if "angle" in want and self.A and "angle" not in res:
            res["angle"] = np.empty((T, self.A), dtype=np.float32)
        if "flags" in want and "flags" not in res:
            res["flags"] = np.empty((T, self.W), dtype=np.uint32)
        if "score" in want and "score" not in res:
            res["score"] = np.empty((T,), dtype=np.int32)
        if "agg" in want:
            res["occupancy"] = np.zeros(self.F, dtype=np.uint64)
            res["sum"] = np.zeros(self.F, dtype=np.float64)
            res["sumsq"] = np.zeros(self.F, dtype=np.float64)
        elif "occ" in want:                          # occupancy without the moments: with no dist / angle -> lean run
            res["occupancy"] = np.zeros(self.F, dtype=np.uint64)
        ho = CvHostOutputs()
        wq = sq = None
        if weights is not None:
            wq = quantize_weights(weights, self.F)
            sq = np.empty((T,), dtype=np.int64)
            ho.weights_q32, ho.score_q32 = wq.ctypes.data, sq.ctypes.data
        shapes = {"dist": ((T, self.P), 4), "angle": ((T, self.A), 4), "flags": ((T, self.W), 4), "score": ((T,), 4),
                  "occupancy": ((self.F,), 8), "sum": ((self.F,), 8), "sumsq": ((self.F,), 8)}
        kinds = {"dist": "f", "angle": "f", "flags": "iu", "score": "i", "occupancy": "iu", "sum": "f", "sumsq": "f"}
        for name in ("dist", "angle", "flags", "score", "occupancy", "sum", "sumsq"):
            arr = res.get(name)
            if arr is not None:
                # raw pointers cross the ABI: a caller-supplied array of the wrong shape / dtype would be overrun
                shape, item = shapes[name]
                if isinstance(arr, np.ndarray):
                    ok = arr.shape == shape and arr.dtype.itemsize == item and arr.dtype.kind in kinds[name] and arr.flags.c_contiguous
                    ptr = arr.ctypes.data
                else:                               # a (pinned) torch CPU tensor
                    ok = tuple(arr.shape) == shape and arr.element_size() == item and arr.is_contiguous() and not arr.is_cuda \
                        and (arr.dtype.is_floating_point == (kinds[name] == "f"))
                    ptr = arr.data_ptr()
                if not ok:
                    raise ValueError(f"output '{name}' must be a C-contiguous host array of shape {shape} with {item}-byte "
                                     f"{'float' if kinds[name] == 'f' else 'integer'} items")
                setattr(ho, name, ptr)
        check(_lib.lib.caviar_features_host(self._handle, xyz.ctypes.data, 3 * n_in,
                                            None if b is None else b.ctypes.data, T, C.byref(ho), int(frames_per_block),
                                            ORDER_PLAN if plan_order else ORDER_CALLER))
        self.last_launches = int(_lib.lib.caviar_last_launch_count())
        if sq is not None:
            res["score_weighted"] = sq.astype(np.float64) / 4294967296.0
        return res


def quantize_weights(weights, n_features: int) -> np.ndarray:
    """Weights -> 32.32 fixed point (int64), the form ``caviar_weighted_scores`` sums exactly.  |w| < 2^20."""
    w = np.asarray(weights, dtype=np.float64).reshape(-1)
    if w.shape[0] != n_features:
        raise ValueError(f"need one weight per feature ({n_features}), got {w.shape[0]}")
    if not np.isfinite(w).all() or np.abs(w).max(initial=0.0) >= 2.0 ** 20:
        raise ValueError("weights must be finite and smaller than 2^20 in magnitude")
    return np.rint(w * 4294967296.0).astype(np.int64)


def unpack_flags(words: np.ndarray, n_pairs: int, n_triplets: int):
    """(T, W) uint32 words -> ((T, P) bool, (T, A) bool).  Pair words come first (include/caviar_b200.h)."""
    w = np.ascontiguousarray(words).view(np.uint32)
    wp = (n_pairs + 31) // 32
    bits = np.unpackbits(w.view(np.uint8).reshape(w.shape[0], -1), axis=1, bitorder="little").astype(bool)
    return bits[:, :n_pairs], bits[:, 32 * wp:32 * wp + n_triplets]
