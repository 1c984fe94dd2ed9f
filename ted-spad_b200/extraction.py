"""Snippet feature-extraction driver: decoded uint8 frames -> crop/resize (CUDA) -> anonymizer UNet
-> 3-D encoder -> one feature row per 16-frame snippet (optionally 5/10-crop) -> `<video>.npy`.

Mirrors the two reference scripts, minus video decoding (frames arrive decoded):
  feature_extraction/dali_extraction.py:103-182   (UCF-Crime / XD-Violence, DALI reader semantics)
  feature_extraction/st_feature_extraction.py:58-100 + shanghai_dl.py:43-98 (ShanghaiTech)
Differences by design (same results, B200-first execution): snippets are batched instead of one
clip per iteration, the per-clip blocking D2H + O(n^2) np.vstack (dali_extraction.py:179) becomes
one D2H per video, and videos are sharded across GPUs by estimated work with no collective.
"""
import os

import numpy as np
import torch

from . import _lib as L
from . import ops


# ------------------------------------------------------------------------------- snippet indexing
def dali_snippet_frames(n_frames, num_frames=16, stride=2):
    """fn.readers.video(sequence_length=16, stride=2, step=32, pad_sequences=True) (dali_extraction.py:58-76):
    snippet i holds frames 32*i + 2*j; a tail snippet is kept and its missing frames are zero images (-1)."""
    step = num_frames * stride
    out = np.full((max(0, -(-n_frames // step)), num_frames), -1, dtype=np.int64)
    for i in range(out.shape[0]):
        idx = i * step + stride * np.arange(num_frames)
        out[i] = np.where(idx < n_frames, idx, -1)
    return out


def shanghai_snippet_frames(n_frames, num_frames=16, fix_skip=2, total_frames=None):
    """shanghai_dl.py:43-98: 1-based frame counter, keep when count % skip == 0, emit a clip when
    count % (16*skip) == 0 (tail dropped); < 32 frames -> skip 1; < 16 frames -> last frame repeated.
    `n_frames` = frames actually decoded; `total_frames` = the container's frame count, which is what the reference
    bases the skip / repeat decisions on (cap.get(7), :50,59-64); None = the same number."""
    total = n_frames if total_frames is None else int(total_frames)
    skip = 1 if total < fix_skip * num_frames else fix_skip
    per = num_frames * skip
    full = n_frames // per
    out = [[i * per + skip * (j + 1) - 1 for j in range(num_frames)] for i in range(full)]
    if total < num_frames and n_frames > 0:
        # repeat: the frame read at count == total is stacked until the clip holds 16 (shanghai_dl.py:84-94)
        if n_frames < total:
            raise RuntimeError(f"only {n_frames} of the {total} frames the container reports could be decoded "
                               "(the reference's keep_frame is never set: the video 'could not process')")
        left = list(range(full * per, n_frames))        # skip == 1 here: every decoded frame is kept
        count = n_frames
        while count % 16 != 0:
            count += 1
            left.append(total - 1)
        if len(left) == num_frames:
            out.append(left)
    return np.asarray(out, dtype=np.int64).reshape(-1, num_frames)


# ------------------------------------------------------------------------------------ crop boxes
def center_offsets(h, w, ch, cw):
    """torchvision center_crop offsets (functional.py:592-593)."""
    return int(round((h - ch) / 2.0)), int(round((w - cw) / 2.0))


def crop_boxes(h, w, ncrops=1, cropping_factor=0.8, no_ar_distortion=False, square_from_h=False):
    """((crop_h, crop_w), [(top, left, hflip), ...]).  Crop size as dali_extraction.py:45-48
    (square_from_h: shanghai_dl.py:35 uses H for both sides); crop order for 5/10 crops is torchvision's
    five_crop / ten_crop: tl, tr, bl, br, center (+ the same five of the h-flipped frame), so that
    crop 4 (center) is the reference's single crop."""
    if no_ar_distortion and square_from_h:
        # shanghai_dl.py:30 takes min(image.shape) of the (H, W, 3) array, i.e. 3: a 2 x 2 crop with the default
        # factor.  Reproduced as written (the reference ships no_ar_distortion=False, params_feature_ex.py:8).
        ch = cw = int(min(h, w, 3) * cropping_factor)
    elif no_ar_distortion:
        ch = cw = int(min(h, w) * cropping_factor)
    elif square_from_h:
        ch = cw = int(h * cropping_factor)
    else:
        ch, cw = int(h * cropping_factor), int(w * cropping_factor)
    ct, cl = center_offsets(h, w, ch, cw)
    if ncrops == 1:
        return (ch, cw), [(ct, cl, 0)]
    five = [(0, 0), (0, w - cw), (h - ch, 0), (h - ch, w - cw), (ct, cl)]
    boxes = [(t, l, 0) for t, l in five]
    if ncrops == 10:
        boxes += [(t, l, 1) for t, l in five]
    elif ncrops != 5:
        raise ValueError("ncrops must be 1, 5 or 10")
    return (ch, cw), boxes


# -------------------------------------------------------------------------------------- extractor
class SnippetExtractor:
    """fa_model / ft_model are the nn.Modules returned by aux_code.model_loaders (already .cuda().eval())."""

    def __init__(self, fa_model, ft_model, reso=(224, 224), num_frames=16, fix_skip=2, cropping_factor=0.8,
                 no_ar_distortion=False, ncrops=1, source="dali", batch_clips=8, device=None):
        self.fa = fa_model
        self.ft = ft_model.i3d if (not hasattr(ft_model, "features_from_cl") and hasattr(ft_model, "i3d")) else ft_model
        self.reso, self.T, self.skip = tuple(reso), num_frames, fix_skip
        self.cf, self.no_ar, self.ncrops = cropping_factor, no_ar_distortion, ncrops
        if source not in ("dali", "shanghai"):
            raise ValueError("source must be 'dali' or 'shanghai'")
        self.source = source
        self.resample = L.RESAMPLE_AA_FLOAT if source == "dali" else L.RESAMPLE_PIL_U8
        self.batch_clips = int(batch_clips)
        self.device = torch.device(device) if device is not None else next(fa_model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("SnippetExtractor needs CUDA modules: there is no CPU path")
        self._enc = {}
        self._copy_stream = None
        self._stage_bufs, self._stage_free = [None, None], [None, None]
        self._desc_ring, self._desc_i = [], 0   # pinned host staging of the per-image descriptors (+ device copy)
        self.h2d_bytes = 0                       # frame bytes staged host -> device so far (bench statistics)
        from .engine import GraphCache
        self._graphs = GraphCache()

    @classmethod
    def from_params(cls, fa_model, ft_model, params, **overrides):
        """Build from a `params_feature_ex`-style module / namespace (feature_extraction/params_feature_ex.py:2-9, read by
        dali_extraction.py:36-50,56-76 and shanghai_dl.py:22-40): num_frames, fix_skip, reso_h, reso_w, cropping_factor,
        no_ar_distortion.  The reference's `batch_size` (1: one clip per DALI batch) and `num_workers` describe ITS loop,
        not the result, and are not taken over - snippets are batched `batch_clips` at a time here, every clip computed
        independently of its neighbours; pass batch_clips / ncrops / source / device as keyword overrides."""
        kw = dict(reso=(int(params.reso_h), int(params.reso_w)), num_frames=int(params.num_frames),
                  fix_skip=int(params.fix_skip), cropping_factor=float(params.cropping_factor),
                  no_ar_distortion=bool(params.no_ar_distortion))
        kw.update(overrides)
        return cls(fa_model, ft_model, **kw)

    def snippet_frames(self, n_frames, total_frames=None):
        if self.source == "dali":
            return dali_snippet_frames(n_frames, self.T, self.skip)
        return shanghai_snippet_frames(n_frames, self.T, self.skip, total_frames)

    def _enc_in(self, B):
        """Encoder input clip [B,T,H,W,4|8]: one allocation for the largest batch seen, smaller batches (video tails)
        run on its B-prefix view, so the footprint does not grow with the number of distinct tail sizes."""
        t = self._enc.get(B)
        if t is None:
            from . import engine
            full = self._enc.get("full")
            if full is None or full.N < B:
                self._enc.clear()
                full = ops.CLTensor(B, self.T, self.reso[0], self.reso[1], engine.ENC_IN_CHANNELS, device=self.device)
                full.buf.zero_()  # pad channels stay zero (the glue never writes them)
                self._enc["full"] = full
            t = full if full.N == B else ops.CLTensor(B, full.D, full.H, full.W, full.C, ld=full.ld, buf=full.buf[:B])
            self._enc[B] = t
        return t

    def features_of_clips(self, frames_dev, desc_host, crop_hw, desc_dev=None):
        """frames_dev: cuda uint8 [F,H,W,3]; desc_host: int32 [B*T,4] numpy -> fp32 cuda [B, n_feat_rows, F].
        desc_dev: the same rows already on the device (features_stream ships them with the frames)."""
        B = desc_host.shape[0] // self.T
        self._check_desc(desc_host, frames_dev.shape, crop_hw)
        with torch.cuda.device(self.device):
            desc = desc_dev if desc_dev is not None else self._desc_to_device(desc_host)
            ex_fa, ex_ft = self.fa.executor(self.device), self.ft.executor(self.device)
            x0 = ex_fa.input_buffer(B * self.T, self.reso[0], self.reso[1])
            ops.preprocess(frames_dev, desc, crop_hw, x0, self.resample)
            enc_in = self._enc_in(B)
            # anonymizer + glue + encoder over fixed buffers: ~100 launches, replayed as ONE CUDA graph per batch size
            gen = (ex_fa.serial, ex_ft.serial, ex_fa.bufs.generation, ex_ft.bufs.generation, x0.buf.data_ptr(), enc_in.buf.data_ptr())
            feats = self._graphs.run((B,), gen, lambda: self.ft.features_from_cl(self.fa.anonymize_into(x0, enc_in, self.T)))
            return feats.clone()   # (the graph's output buffer is overwritten by the next batch)

    @staticmethod
    def _check_desc(desc, frames_shape, crop_hw):
        """Host-side validation of the (src_frame, top, left, hflip) rows: the preprocessing kernel indexes the frame
        buffer with them unchecked."""
        F_, Hs, Ws = int(frames_shape[0]), int(frames_shape[1]), int(frames_shape[2])
        d = np.asarray(desc)
        if d.ndim != 2 or d.shape[1] != 4 or d.dtype != np.int32:
            raise ValueError("desc must be int32 [n, 4] = (src_frame, top, left, hflip)")
        if d.size and (int(d[:, 0].max()) >= F_ or int(d[:, 1].min()) < 0 or int(d[:, 2].min()) < 0 or
                       int(d[:, 1].max()) + int(crop_hw[0]) > Hs or int(d[:, 2].max()) + int(crop_hw[1]) > Ws):
            raise ValueError(f"desc addresses pixels outside the [{F_},{Hs},{Ws}] frame buffer "
                             f"(crop {tuple(crop_hw)}, max frame {int(d[:, 0].max())}, max top {int(d[:, 1].max())}, "
                             f"max left {int(d[:, 2].max())})")

    def _desc_to_device(self, desc_host):
        """Per-image descriptors through a small ring of PINNED host buffers (a pageable source makes the copy
        synchronous with the host) into per-slot device buffers; a slot is reused only after its copy has run."""
        n = desc_host.shape[0]
        if not self._desc_ring:
            self._desc_ring = [[None, None, None] for _ in range(4)]
        slot = self._desc_ring[self._desc_i % 4]
        self._desc_i += 1
        if slot[0] is None or slot[0].shape[0] < n:
            if slot[2] is not None:
                slot[2].synchronize()
            slot[0] = torch.empty((max(n, 512), 4), dtype=torch.int32).pin_memory()
            slot[1] = torch.empty((max(n, 512), 4), dtype=torch.int32, device=self.device)
        elif slot[2] is not None:
            slot[2].synchronize()
        slot[0][:n].numpy()[...] = desc_host
        dev = slot[1][:n]
        dev.copy_(slot[0][:n], non_blocking=True)
        slot[2] = torch.cuda.Event()
        slot[2].record()
        return dev

    def _stage(self, frames, slot, desc_host=None):
        """Host frames -> one of two persistent device staging buffers, on the copy stream.  Returns the device view,
        the device copy of `desc_host` (shipped on the SAME stream right behind the frames: a pinned 8 KB copy issued
        on the compute stream would queue on the copy engine behind the NEXT batch's 118 MB transfer and stall the
        step by 2-5 ms) and the event that marks their arrival.  (Fresh allocations per batch made the caching allocator cudaMalloc -
        a device-wide synchronisation - whenever the previous batch's block was still in use.)"""
        chunks = list(frames) if isinstance(frames, (list, tuple)) else [frames]
        if all(c.is_cuda for c in chunks):
            return (chunks[0] if len(chunks) == 1 else torch.cat(chunks, 0)).contiguous(), None, None
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        chunks = [c.contiguous() for c in chunks]
        n = sum(c.numel() for c in chunks)
        self.h2d_bytes += n
        buf = self._stage_bufs[slot]
        if buf is None or buf.numel() < n:
            # allocated ON the copy stream (the stream that writes it); a block the caching allocator hands back from
            # the compute stream could still be read by kernels in flight there.  The buffer being replaced may still
            # be read by the compute stream: tell the allocator before dropping it.
            if buf is not None:
                buf.record_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._copy_stream):
                buf = torch.empty(max(n, 1), dtype=torch.uint8, device=self.device)
            self._stage_bufs[slot] = buf
        dev = buf[:n].view((sum(c.shape[0] for c in chunks),) + tuple(chunks[0].shape[1:]))
        with torch.cuda.stream(self._copy_stream):
            if self._stage_free[slot] is not None:
                self._copy_stream.wait_event(self._stage_free[slot])   # the batch that used this buffer has been consumed
            f0 = 0
            for c in chunks:   # frame ranges of one or several videos, back to back in the staging buffer
                dev[f0:f0 + c.shape[0]].copy_(c, non_blocking=True)
                f0 += c.shape[0]
            desc_dev = self._desc_to_device(desc_host) if desc_host is not None else None
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return dev, desc_dev, ev

    def features_stream(self, batches):
        """batches: iterable of (frames uint8 [F,H,W,3] on the host (pinned for a truly asynchronous copy) or on the
        device, desc int32 [B*T,4] numpy, crop_hw).  Yields fp32 cuda features [B, n_feat_rows, F] per batch.  The
        host->device copy of batch i+1 runs on a copy stream while batch i is computed (the reference hands DALI's
        GPU tensors over one clip at a time, dali_extraction.py:147-150)."""
        it = iter(batches)
        nxt = next(it, None)
        i = 0
        staged = self._stage(nxt[0], 0, nxt[1]) if nxt is not None else None
        while nxt is not None:
            cur, (dev, desc_dev, ev) = nxt, staged
            nxt = next(it, None)
            if nxt is not None:
                staged = self._stage(nxt[0], (i + 1) % 2, nxt[1])  # in flight while `cur` is computed
            compute = torch.cuda.current_stream(self.device)
            if ev is not None:
                compute.wait_event(ev)
            feats = self.features_of_clips(dev, cur[1], cur[2], desc_dev)
            if ev is not None:
                done = torch.cuda.Event()
                done.record(compute)
                self._stage_free[i % 2] = done                      # buffer i % 2 may be overwritten after this point
            i += 1
            yield feats

    SPARSE_FRAME_BYTES = 400 * 1024   # per-frame copies pay off above this frame size (a copy call costs ~10 us)

    def _needed_frames(self, frames, sn, base):
        """Frames of `frames` a group of snippets `sn` (int64 [n, T], -1 = zero image) reads -> (chunks to stage, the
        snippets' indices into the staged frames starting at `base`, number of staged frames).  Small frames: the
        contiguous range [min, max] in one copy.  Large frames (480x856 ShanghaiTech: 1.2 MB each, of which the reader
        keeps every second one, shanghai_dl.py:73): only the frames that are used, one copy per run of consecutive
        frames - 20 instead of 38 MB of host->device traffic per clip, which at ~1000 clips/s is the difference between
        fitting the PCIe link and saturating it."""
        valid = sn[sn >= 0]
        if not valid.size:
            return [frames[0:1]], np.full(sn.shape, -1, dtype=np.int64), 1
        f_lo, f_hi = int(valid.min()), int(valid.max()) + 1
        frame_bytes = int(frames.shape[1]) * int(frames.shape[2]) * 3
        uniq = np.unique(valid)
        if frame_bytes < self.SPARSE_FRAME_BYTES or uniq.size > 0.75 * (f_hi - f_lo) or frames.is_cuda:
            return [frames[f_lo:f_hi]], np.where(sn >= 0, sn - f_lo + base, -1), f_hi - f_lo
        starts = np.flatnonzero(np.diff(uniq, prepend=uniq[0] - 2) != 1)          # first frame of every run
        ends = np.append(starts[1:], uniq.size)
        chunks = [frames[int(uniq[a]):int(uniq[b - 1]) + 1] for a, b in zip(starts, ends)]
        pos = np.searchsorted(uniq, np.where(sn >= 0, sn, uniq[0]))
        return chunks, np.where(sn >= 0, pos + base, -1), int(uniq.size)

    def extract_videos(self, videos):
        """videos: iterable of uint8 [F,H,W,3] tensors or of callables returning them.  Yields (index, features) in
        input order, features as extract_video returns them.  Snippets are packed into full batches ACROSS videos
        (frames of equal size share a batch): datasets of short videos - ShanghaiTech averages ~25 snippets per video -
        would otherwise run one partial batch per video.  Every clip is computed independently of its batch
        neighbours, so the rows are bit-identical to extract_video's (tests/test_gpu_parity.py)."""
        import collections
        per_batch = max(1, self.batch_clips // self.ncrops)   # snippets per batch
        recs = collections.deque()      # videos in flight, in input order
        metas = collections.deque()     # per issued batch: [(record, snippets taken), ...]

        def batches():
            chunks, descs, meta, count, base, hw, crop = [], [], [], 0, 0, None, None

            def flush():
                nonlocal chunks, descs, meta, count, base
                out = (chunks, np.concatenate(descs, 0).reshape(-1, 4), crop)
                metas.append(meta)
                chunks, descs, meta, count, base = [], [], [], 0, 0
                return out

            for idx, v in enumerate(videos):
                frames = v() if callable(v) else v
                n_frames, H, W, _ = frames.shape
                snips = self.snippet_frames(n_frames)
                crop_hw, boxes = crop_boxes(H, W, self.ncrops, self.cf, self.no_ar, square_from_h=(self.source == "shanghai"))
                rec = {"idx": idx, "left": snips.shape[0], "rows": []}
                recs.append(rec)
                if count and (H, W) != hw:
                    yield flush()
                hw, crop = (H, W), crop_hw
                s0 = 0
                while s0 < snips.shape[0]:
                    take = min(per_batch - count, snips.shape[0] - s0)
                    sn = snips[s0:s0 + take]
                    parts, rel, n_staged = self._needed_frames(frames, sn, base)
                    desc = np.empty((take, len(boxes), self.T, 4), dtype=np.int32)
                    desc[..., 0] = rel[:, None, :]
                    for ci, (t, l, fl) in enumerate(boxes):
                        desc[:, ci, :, 1], desc[:, ci, :, 2], desc[:, ci, :, 3] = t, l, fl
                    chunks.extend(parts)
                    descs.append(desc.reshape(-1, 4))
                    meta.append((rec, take))
                    base += n_staged
                    count += take
                    s0 += take
                    if count == per_batch:
                        yield flush()
            if count:
                yield flush()

        pending = collections.deque()   # videos whose rows are on their way to the host: (index, pinned rows, event)

        def finished(flush=False):
            """Completed videos, in input order.  Their rows leave the device with an asynchronous copy into pinned
            memory and are handed out one or two batches later, when the copy has landed: a blocking read-back here
            would stall the host until the GPU has finished the batch it has only just been given, and the GPU would
            then idle while the next batch is prepared (measured: 4 % of a 10-crop dataset run)."""
            while recs and recs[0]["left"] == 0:
                rec = recs.popleft()
                if not rec["rows"]:
                    pending.append((rec["idx"], None, None))
                    continue
                allf = torch.cat(rec["rows"], 0)
                if not allf.is_cuda:
                    pending.append((rec["idx"], allf, None))
                    continue
                host = torch.empty(allf.shape, dtype=allf.dtype, pin_memory=True)
                host.copy_(allf, non_blocking=True)                    # one D2H per video
                ev = torch.cuda.Event()
                ev.record()
                pending.append((rec["idx"], host, ev))
            while pending and (flush or len(pending) > 2 or pending[0][2] is None or pending[0][2].query()):
                idx, host, ev = pending.popleft()
                if host is None:
                    yield idx, np.zeros((0, 0), dtype=np.float64)
                    continue
                if ev is not None:
                    ev.synchronize()
                allf = host.numpy().astype(np.float64)
                yield idx, (allf[:, 0, :] if self.ncrops == 1 else allf)

        for f in self.features_stream(batches()):
            f = f.reshape(-1, self.ncrops, f.shape[-1] * f.shape[-2])
            r0 = 0
            for rec, take in metas.popleft():
                rec["rows"].append(f[r0:r0 + take])
                rec["left"] -= take
                r0 += take
            yield from finished()
        yield from finished(flush=True)

    def extract_video(self, frames):
        """frames: uint8 [F,H,W,3] torch tensor (CPU, pinned CPU or CUDA), RGB for the DALI path, BGR
        as cv2 decodes for the ShanghaiTech path.  Returns float64 numpy [n_snip, feat] (ncrops == 1, the
        reference layout: dali_extraction.py:163,179) or [n_snip, ncrops, feat] (dataset.py:70-71,89)."""
        n_frames, H, W, _ = frames.shape
        snips = self.snippet_frames(n_frames)
        n_snip = snips.shape[0]
        crop_hw, boxes = crop_boxes(H, W, self.ncrops, self.cf, self.no_ar, square_from_h=(self.source == "shanghai"))
        per_batch = max(1, self.batch_clips // self.ncrops)  # snippets per batch

        def batches():
            for s0 in range(0, n_snip, per_batch):
                sn = snips[s0:s0 + per_batch]
                parts, rel, _ = self._needed_frames(frames, sn, 0)
                desc = np.empty((sn.shape[0], len(boxes), self.T, 4), dtype=np.int32)
                desc[..., 0] = rel[:, None, :]
                for ci, (t, l, fl) in enumerate(boxes):
                    desc[:, ci, :, 1], desc[:, ci, :, 2], desc[:, ci, :, 3] = t, l, fl
                yield parts, desc.reshape(-1, 4), crop_hw

        rows = [f.reshape(-1, len(boxes), f.shape[-1] * f.shape[-2]).clone() for f in self.features_stream(batches())]
        if not rows:
            width = 0
            return np.zeros((0, width), dtype=np.float64)
        allf = torch.cat(rows, 0).cpu().numpy().astype(np.float64)  # one D2H per video
        return allf[:, 0, :] if self.ncrops == 1 else allf


def segment_features(vid_features, num_features=None, rule="dali"):
    """The optional 32-segment pooling of the reference scripts ("as in Sultani et al."), which both scripts ship
    DISABLED (`segment=False`: dali_extraction.py:155,181; st_feature_extraction.py:100).  Restated as written:
    segment boundaries `np.linspace(0, n, 33, dtype=int)`; the DALI script's branch `ss <= es or es < ss` is always
    true (dali_extraction.py:93), so every segment is the single row at its start, L2-normalised; the ShanghaiTech
    script (st_feature_extraction.py:49-52) takes that row when `ss <= es` and otherwise the *scalar*
    `np.mean(vid_features[ss:es])` (NaN for an empty slice, the mean of all rows but the last when es == -1)
    divided by its own norm and broadcast over the row.  Returns
    float64 [32, F]."""
    vid_features = np.asarray(vid_features)
    n, F = vid_features.shape
    if num_features is not None and num_features != F:
        raise ValueError(f"features have {F} columns, expected {num_features}")
    if rule not in ("dali", "shanghai"):
        raise ValueError("rule must be 'dali' or 'shanghai'")
    loc = np.linspace(0, n, 33, dtype=int)
    out = np.zeros((32, F))
    with np.errstate(invalid="ignore", divide="ignore"):
        for idx in range(32):
            ss, es = loc[idx], loc[idx + 1] - 1
            if idx == 31:
                es += 1
            if rule == "dali" or ss <= es:
                v = vid_features[ss]                          # raises for an empty video, like the reference
            else:
                v = np.mean(vid_features[ss:es])              # scalar; Python slice semantics (es may be -1)
            out[idx] = v / np.linalg.norm(v)
    return out


def feature_path(save_folder, vid_path):
    """<folder>/<basename minus .mp4/.avi>.npy (dali_extraction.py:159, st_feature_extraction.py:88)."""
    base = os.path.basename(vid_path).replace('.mp4', '').replace('.avi', '')
    return os.path.join(save_folder, base + '.npy')


# ------------------------------------------------------------------------------------- sharding
def shard_videos(lengths, world_size):
    """Longest-processing-time greedy partition of videos over ranks by frame count.
    Returns a list (per rank) of video indices; every video is owned by exactly one rank."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(lengths[i])
    return [sorted(s) for s in shards]


def extract_dataset(extractor, videos, save_folder, rank=0, world_size=1, log=print, segment=False, prefetch_workers=0):
    """videos: list of (path, n_frames, loader) with loader() -> uint8 [F,H,W,3].  prefetch_workers > 0: the loaders of
    the next videos run on that many threads while the current one is extracted (ingest.prefetched; for file decoders).  The partition is computed
    over the FULL list (so every rank derives the same one no matter when it starts); within its shard a
    rank skips videos whose .npy already exists - the reference's resume rule (dali_extraction.py:121).
    Each file is written by exactly one rank (atomic rename); there is no collective."""
    os.makedirs(save_folder, exist_ok=True)
    mine = shard_videos([v[1] for v in videos], world_size)[rank]
    todo = [i for i in mine if not os.path.exists(feature_path(save_folder, videos[i][0]))]
    written = []

    def save(i, feats):
        out = feature_path(save_folder, videos[i][0])
        if segment:   # the reference's `segment` flag (False in both scripts: dali_extraction.py:181, st:100)
            feats = segment_features(feats, rule="shanghai" if getattr(extractor, "source", "dali") == "shanghai" else "dali")
        tmp = out + f".tmp{rank}.npy"
        np.save(tmp, feats)
        os.replace(tmp, out)
        written.append(out)

    def load(i):
        log(f'Extracting features for {os.path.basename(videos[i][0])}.')
        return videos[i][2]()

    if hasattr(extractor, "extract_videos"):
        # snippets packed into full batches across this rank's videos (same rows, see SnippetExtractor.extract_videos)
        sources = ((lambda i=i: load(i)) for i in todo)
        if prefetch_workers > 0:
            from .ingest import prefetched
            sources = prefetched(sources, workers=prefetch_workers)      # decoded tensors instead of callables
        for k, feats in extractor.extract_videos(sources):
            save(todo[k], feats)
    else:
        for i in todo:
            save(i, extractor.extract_video(load(i)))
    return written


def extract_dataset_distributed(extractor, videos, save_folder, log=print, stats=None):
    """extract_dataset under torch.distributed (one process per GPU, `torchrun`): rank / world size come from the
    process group, every rank extracts its own shard with NO collective on the data path, and the run manifest
    (which rank wrote which file) is gathered on the host afterwards - the "host-side gather of feature files" of
    the north star.  Returns {path: rank} on every rank.  Works with the gloo and nccl backends."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return {w: 0 for w in extract_dataset(extractor, videos, save_folder, 0, 1, log)}
    rank, world = dist.get_rank(), dist.get_world_size()
    import time
    t0 = time.perf_counter()
    written, err = [], None
    try:
        written = extract_dataset(extractor, videos, save_folder, rank, world, log)
    except Exception as e:  # noqa: BLE001  (reported on EVERY rank below: a rank that dies alone leaves the others
        err = f"rank {rank}: {type(e).__name__}: {e}"          # waiting in the gather forever)
    if stats is not None:   # this rank's own shard, before the closing manifest gather (which waits for the slowest rank)
        if torch.cuda.is_available() and err is None:
            torch.cuda.synchronize()
        stats["extract_s"] = time.perf_counter() - t0
    gathered = [None] * world
    dist.all_gather_object(gathered, (written, err))   # control plane only: file names, after the work is done
    errors = [e for _, e in gathered if e]
    if errors:
        raise RuntimeError("sharded extraction failed: " + "; ".join(errors))
    gathered = [w for w, _ in gathered]
    manifest = {}
    for r, files in enumerate(gathered):
        for f in files:
            if f in manifest:
                raise RuntimeError(f"{f} was written by ranks {manifest[f]} and {r}: the shards overlap")
            manifest[f] = r
    return manifest
