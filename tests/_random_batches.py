"""Random mixtures of every input modality for the planner / tokeniser tests."""
import numpy as np
import torch


def random_mixed_batch(seed: int):
    """Returns (batch, context_len, pad_seq): text lists / tensors, uint8 / float frames of several sizes, continuous /
    discrete observations and actions, ragged lengths."""
    rs = np.random.RandomState(1000 + seed)
    f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))  # noqa: E731
    ctx = 160
    pad_seq = bool(seed % 2)
    batch = []
    for _ in range(int(rs.randint(1, 7))):
        kind = rs.randint(0, 6)
        if kind == 0:
            n = int(rs.randint(1, 40))
            ids = rs.randint(0, 300, (n,))
            batch.append({"text": ids.tolist() if rs.rand() < 0.5 else torch.from_numpy(ids)})
        elif kind == 1:
            T, o, a = int(rs.randint(1, 5)), int(rs.randint(1, 9)), int(rs.randint(1, 5))
            batch.append({"continuous_obs": f32(rs.standard_normal((T, o)) * 3), "continuous_actions": f32(np.clip(rs.standard_normal((T, a)), -1, 1))})
        elif kind == 2:
            T = int(rs.randint(1, 3))
            h, w = 16 * int(rs.randint(1, 4)), 16 * int(rs.randint(1, 3))
            img = rs.randint(0, 256, (T, 3, h, w))
            batch.append({"images": torch.from_numpy(img.astype(np.uint8)) if rs.rand() < 0.5 else f32(img),
                          "discrete_actions": torch.from_numpy(rs.randint(0, 5, (T, 1)).astype(np.int32))})
        elif kind == 3:
            T, do = int(rs.randint(1, 5)), int(rs.randint(1, 6))
            batch.append({"discrete_obs": torch.from_numpy(rs.randint(0, 9, (T, do)).astype(np.int64)),
                          "continuous_actions": f32(np.clip(rs.standard_normal((T, 2)), -1, 1))})
        elif kind == 4:
            img = rs.randint(0, 256, (1, 3, 32, 32)).astype(np.uint8)
            batch.append({"images": torch.from_numpy(img), "text": torch.from_numpy(rs.randint(0, 300, (int(rs.randint(1, 20)),)))})
        else:
            T, o = int(rs.randint(1, 4)), int(rs.randint(1, 6))
            batch.append({"continuous_obs": f32(rs.standard_normal((T, o))), "discrete_actions": torch.from_numpy(rs.randint(0, 4, (T, 1)).astype(np.int32))})
    return batch, ctx, pad_seq
