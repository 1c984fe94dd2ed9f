"""ORACLE (test infrastructure, never imported by the product package).

numpy restatement of how the downstream consumer reads the feature files this path writes:

  load_features      anomaly_detection_mgfn/datasets/dataset.py:53-55  np.load(..., allow_pickle=True) -> float32
  getitem_test       dataset.py:68-86    [T,F] -> [T,1,F] (expand_dims axis 1), L2 magnitude appended as feature F+1
  getitem_train      dataset.py:87-99    -> transpose to [ncrops,T,F], every crop resampled to `seg_length` segments
                                          by process_feat, magnitude of the *segment* rows appended -> [ncrops,32,F+1]
  process_feat       anomaly_detection_mgfn/utils/utils.py:34-42

(The three `datasetname` branches of dataset.py are textually identical; one restatement covers UCF / XD / ST.)
Pinned against the reference's own Dataset.__getitem__ / process_feat run in this container:
tests/golden/make_golden_consumer.py -> tests/golden/consumer_v1.npz, checked by tests/test_oracle.py.
"""
import numpy as np


def load_features(path):
    """dataset.py:53-55."""
    features = np.load(path, allow_pickle=True)
    return np.array(features, dtype=np.float32)


def process_feat(feat, length=32):
    """utils/utils.py:34-42: `length` segments with boundaries linspace(0, T, length+1, dtype=int); a segment is the
    mean of its rows, an EMPTY segment (T < length) takes the single row at its start index."""
    feat = np.asarray(feat)
    out = np.zeros((length, feat.shape[1]), dtype=np.float32)
    r = np.linspace(0, len(feat), length + 1, dtype=int)
    for i in range(length):
        if r[i] != r[i + 1]:
            out[i, :] = np.mean(feat[r[i]:r[i + 1], :], 0)
        else:
            out[i, :] = feat[r[i], :]
    return out


def getitem_test(features):
    """dataset.py:68-86 -> float32 [T, ncrops, F+1]."""
    features = np.array(features, dtype=np.float32)
    if features.ndim < 3:
        features = np.expand_dims(features, axis=1)
    mag = np.linalg.norm(features, axis=2)[:, :, np.newaxis]
    return np.concatenate((features, mag), axis=2)


def getitem_train(features, seg_length=32):
    """dataset.py:87-99 -> float32 [ncrops, seg_length, F+1]."""
    features = np.array(features, dtype=np.float32)
    if features.ndim < 3:
        features = np.expand_dims(features, axis=1)
    features = features.transpose(1, 0, 2)
    divided, mags = [], []
    for feature in features:
        feature = process_feat(feature, seg_length)
        divided.append(feature)
        mags.append(np.linalg.norm(feature, axis=1)[:, np.newaxis])
    divided = np.array(divided, dtype=np.float32)
    mags = np.array(mags, dtype=np.float32)
    return np.concatenate((divided, mags), axis=2)
