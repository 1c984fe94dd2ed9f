"""Video ingest for the extraction driver, host side.

`decode_video_cv2` is the decoding half of the reference's ShanghaiTech reader
(feature_extraction/shanghai_dl.py:43-98): cv2.VideoCapture, every frame decoded in order until `read()` fails,
frames kept in the BGR order cv2 delivers (the reference never converts, shanghai_dl.py:66-75).  The reference then
augments frame by frame on the CPU with PIL; here the decoded uint8 frames go to the GPU as they are and the snippet
selection / crop / Pillow-exact resize / anonymizer / encoder run there (extraction.SnippetExtractor(source="shanghai")).

The DALI reader of dali_extraction.py:53-81 decodes on NVDEC.  Neither `nvidia.dali` nor the Video Codec SDK
(nvcuvid.h / libnvcuvid) is present in this image, so a device-side decoder cannot be built or tested here: the DALI
path takes decoded frames (its reader semantics - stride 2, step 32, zero-padded tail - are in
extraction.dali_snippet_frames), and `decode_video_cv2` can feed it too (RGB conversion on request).
"""
import glob
import os

import numpy as np
import torch


def decode_video_cv2(path, rgb=False, pin=True):
    """-> (frames uint8 [F,H,W,3] torch tensor (pinned when CUDA is available), total_frames as the container reports
    it: cap.get(CAP_PROP_FRAME_COUNT), which shanghai_dl.py:50,59-64 uses for the skip / repeat decisions).
    Raises RuntimeError when the file cannot be opened or holds no frame (the reference returns (None, None, path) from
    a bare except and the caller skips the video, st_feature_extraction.py:93-98)."""
    import cv2
    cap = cv2.VideoCapture(path)
    if not cap.isOpened():
        raise RuntimeError(f"cv2 cannot open {path}")
    cap.set(cv2.CAP_PROP_FPS, 25)          # shanghai_dl.py:47-48 (no effect on file captures; kept for fidelity)
    cap.set(1, 0)
    total = int(cap.get(7))
    frames = []
    while cap.isOpened():
        ret, frame = cap.read()
        if not ret:
            break
        frames.append(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB) if rgb else frame)
    cap.release()
    if not frames:
        raise RuntimeError(f"{path}: no frame could be decoded")
    out = torch.from_numpy(np.stack(frames))
    if pin and torch.cuda.is_available():
        out = out.pin_memory()
    return out, total


def shanghai_video_list(root, reverse=False):
    """shanghai_dl.py:16-19: sorted(glob(<root>/t*/videos/*))."""
    vids = sorted(glob.glob(os.path.join(root, 't*', 'videos', '*')))
    return vids[::-1] if reverse else vids


def cv2_dataset(paths, rgb=False):
    """[(path, n_frames, loader)] for extraction.extract_dataset: n_frames is the container's frame count (used only to
    balance the shards), the loader decodes on demand."""
    import cv2
    out = []
    for p in paths:
        cap = cv2.VideoCapture(p)
        n = int(cap.get(7)) if cap.isOpened() else 0
        cap.release()
        out.append((p, n, (lambda p=p: decode_video_cv2(p, rgb)[0])))
    return out


def prefetched(loaders, workers=4, depth=None):
    """Runs `loaders` (callables returning a decoded video) on a pool of `workers` threads, at most `depth` videos ahead
    of the consumer, and yields their results IN ORDER.  cv2 releases the GIL while it decodes, so the next videos are
    decoded while the GPU works on the current one (the reference decodes and augments on the main thread between two
    videos: st_feature_extraction.py:87, shanghai_dl.py:43-98).  A loader's exception is re-raised at its position."""
    import collections
    from concurrent.futures import ThreadPoolExecutor
    depth = depth or 2 * workers
    it = iter(loaders)
    pending = collections.deque()
    with ThreadPoolExecutor(max_workers=workers) as pool:
        def fill():
            while len(pending) < depth:
                fn = next(it, None)
                if fn is None:
                    return
                pending.append(pool.submit(fn))
        fill()
        while pending:
            fut = pending.popleft()
            res = fut.result()
            fill()
            yield res
