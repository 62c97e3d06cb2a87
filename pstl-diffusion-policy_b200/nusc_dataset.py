"""Offline data path of the reference (``--offline``: everything after ``--collect_data``): the scene cache
``cache.npz`` written by ``nusc_train.py:193-208`` (``np.savez(data={traj_i: {ti: {key: array}}}, meta_list=...)``)
and the per-sample traj-opt files read back by ``nusc_dataset.py:109-118, 203-240``.  The online path (NuScenes devkit
queries, ``nusc_api``) is out of scope; a cache produced by the reference loads here unchanged, and a cache written
here loads in the reference.
"""
import os

import numpy as np
import torch

KEEP_KEYS = ("traj_i", "ti", "len_full")  # left as they are by dict_to_torch (utils.py:72-79)
TRAJOPT_KEYS = ("params", "params_init", "pre_stlp", "tj_scores_prior")


def save_cache_data(batch, saved_sample_d):
    """reference nusc_train.py:193-204: file one collated batch under saved[traj_i][ti][key] (``params`` is not cached)."""
    batch_np = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in batch.items()}
    for i in range(batch_np["traj_i"].shape[0]):
        traj_i, ti = int(batch_np["traj_i"][i]), int(batch_np["ti"][i])
        saved_sample_d.setdefault(traj_i, {})[ti] = {k: v[i] for k, v in batch_np.items() if k != "params"}
    return saved_sample_d


def write_cache(path, saved_sample_d, meta_list):
    """reference nusc_train.py:208."""
    np.savez(path, data=saved_sample_d, meta_list=np.asarray(meta_list, dtype=object))


def read_cache(path):
    """reference nusc_train.py:156-157: (data dict, meta_list)."""
    z = np.load(path, allow_pickle=True)
    return z["data"].item(), z["meta_list"]


def read_split_file(path, test_t1=False):
    """``data/*_split.txt`` lines ``traj_i ti token`` (nusc_dataset.py:84-95)."""
    out = []
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            traj_i, ti, token = line.strip().split(" ")
            if test_t1 and int(ti) != 1:
                continue
            out.append((int(traj_i), int(ti), token))
    return out


class CacheDataset(torch.utils.data.Dataset):
    """``MyDataset`` in ``--offline`` mode: samples come from the cache, the traj-opt parameters from
    ``<params_dir>/params_<traj>_<ti>*.npy`` (random initial controls when there are none, as upstream :214-218),
    re-sampled to ``args.n_randoms`` rows when the stored count differs (:233-240)."""

    def __init__(self, cache, args, indices=None, params_dir=None):
        self.cache = cache
        self.args = args
        if indices is None:
            indices = [(t, k, None) for t in sorted(cache) for k in sorted(cache[t])]
        self.indices = [tuple(ix) if len(ix) == 3 else (ix[0], ix[1], None) for ix in indices]
        self.params_dir = params_dir

    def __len__(self):
        return len(self.indices)

    def __getitem__(self, idx):
        traj_i, ti, _ = self.indices[idx]
        a = self.args
        d = {}
        for k, v in self.cache[traj_i][ti].items():
            if k in TRAJOPT_KEYS:
                continue  # files / fresh draws below, like upstream
            d[k] = v if k in KEEP_KEYS else torch.from_numpy(np.asarray(v)).float()
        key = (int(traj_i), int(ti))
        p = self.params_dir and os.path.join(self.params_dir, "params_%05d_%04d.npy" % key)
        if p and os.path.exists(p):
            d["params"] = torch.from_numpy(np.load(p)).float()
            d["params_init"] = torch.from_numpy(np.load(os.path.join(self.params_dir, "params_%05d_%04d_init.npy" % key))).float()
        else:
            w = (torch.rand(a.n_randoms, 3, a.nt) * 2 - 1) * a.mul_w_max * 0.1
            acc = (torch.rand(a.n_randoms, 3, a.nt) * 2 - 1) * a.mul_a_max
            d["params"] = torch.stack([w, acc], dim=-1)
            d["params_init"] = d["params"].clone()
        if getattr(a, "load_stlp", False) and self.params_dir:
            d["pre_stlp"] = torch.from_numpy(np.load(os.path.join(self.params_dir, "params_%05d_%04d_stlp.npy" % key))).float()
            d["tj_scores_prior"] = torch.from_numpy(np.load(os.path.join(self.params_dir, "scores_%05d_%04d.npy" % key))).float()
        n0 = d["params_init"].shape[0]
        if n0 != a.n_randoms:
            pick = torch.from_numpy(np.random.choice(n0, a.n_randoms))
            for k in TRAJOPT_KEYS:
                if k in d:
                    d[k] = d[k][pick]
        return d


def get_dataloader(args, cache_path, split_file=None, params_dir=None, shuffle=True):
    """``get_dataloader`` for ``--offline`` (reference nusc_train.py:153-191): a torch DataLoader of scene batches."""
    cache, _ = read_cache(cache_path)
    indices = read_split_file(split_file, getattr(args, "test_t1", False)) if split_file else None
    ds = CacheDataset(cache, args, indices, params_dir)
    return torch.utils.data.DataLoader(ds, batch_size=args.batch_size, shuffle=shuffle, num_workers=args.num_workers,
                                       pin_memory=True, drop_last=False)
