"""Latent wire format of the Counter-Strike diffusion data: the reference stores VAE latents as MosaicML
Streaming "MDS" shards written by `MDSWriter(columns={'mean': 'ndarray', 'action': 'ndarray'}, compression='zstd')`
(edm2/cs_dataset_processing/dataset_processing_counter_strike.py:86-98) and reads them through `StreamingDataset`, slicing each
~1000-frame episode into `clip_size` clips (edm2/cs_dataloading.py:53-71) that the training loop normalises with the VAE's
latent statistics (cs_train.py:102).

`mosaicml-streaming` (un-pinned in the reference's pyproject.toml:26) and zstd are not installed in this image, so this
module restates the PUBLISHED MDS v2 layout for *uncompressed* shards and is pinned only by its own round trip -- byte
compatibility with the library is **parity unpinned** (no fixture or library to check against here):

    shard  := uint32 n_samples | uint32 offsets[n_samples + 1] (absolute) | config JSON (utf-8) | samples
    sample := uint32 size[k] for every variable-size column | column payloads, in column order
    ndarray payload (dtype and shape not fixed in the schema) := uint8 dtype code | uint8 ndim | uint32 dims[ndim] | raw data
    index.json := {"version": 2, "shards": [{column_encodings, column_names, column_sizes, compression, format: "mds",
                   hashes, raw_data: {basename, bytes, hashes}, samples, size_limit, version, zip_data}]}

What the training path needs from it is the reader side: `LatentClips` iterates (mean [clip, C, h, w], action [clip, ...])
clips exactly as CsVaeDataset does, and `normalize_latents` is cs_train.py:102 on the device.
"""
import json
import os

import numpy as np
import torch

_DTYPES = [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.uint64, np.int64, np.float16, np.float32, np.float64]


def _encode_ndarray(a):
    a = np.ascontiguousarray(a)
    code = [np.dtype(d) for d in _DTYPES].index(a.dtype)
    return bytes([code, a.ndim]) + np.asarray(a.shape, dtype=np.uint32).tobytes() + a.tobytes()


def _decode_ndarray(buf):
    code, ndim = buf[0], buf[1]
    shape = np.frombuffer(buf, dtype=np.uint32, count=ndim, offset=2)
    return np.frombuffer(buf, dtype=_DTYPES[code], offset=2 + 4 * ndim).reshape(tuple(int(s) for s in shape))


class ShardWriter:
    """Writes samples {'mean': fp16 [c, t, h, w], 'action': ndarray} into uncompressed MDS-layout shards + index.json."""

    def __init__(self, out_dir, columns=("mean", "action"), size_limit=1 << 26):
        self.out_dir, self.columns, self.size_limit = out_dir, list(columns), size_limit
        os.makedirs(out_dir, exist_ok=True)
        self.shards, self.samples, self.bytes = [], [], 0

    def write(self, sample):
        payloads = [_encode_ndarray(sample[c]) for c in self.columns]
        data = np.asarray([len(p) for p in payloads], dtype=np.uint32).tobytes() + b"".join(payloads)
        if self.samples and self.bytes + len(data) > self.size_limit:
            self._flush()
        self.samples.append(data)
        self.bytes += len(data)

    def _flush(self):
        if not self.samples:
            return
        cfg = {"column_encodings": ["ndarray"] * len(self.columns), "column_names": self.columns,
               "column_sizes": [None] * len(self.columns), "compression": None, "format": "mds", "hashes": [],
               "size_limit": self.size_limit, "version": 2}
        cfg_b = json.dumps(cfg, sort_keys=True).encode("utf-8")
        n = len(self.samples)
        offsets = np.zeros(n + 1, dtype=np.uint32)
        offsets[0] = 4 + 4 * (n + 1) + len(cfg_b)
        offsets[1:] = offsets[0] + np.cumsum([len(s) for s in self.samples], dtype=np.uint64)
        name = f"shard.{len(self.shards):05d}.mds"
        with open(os.path.join(self.out_dir, name), "wb") as f:
            f.write(np.uint32(n).tobytes() + offsets.tobytes() + cfg_b + b"".join(self.samples))
        self.shards.append(dict(cfg, raw_data={"basename": name, "bytes": int(offsets[-1]), "hashes": {}}, samples=n, zip_data=None))
        self.samples, self.bytes = [], 0

    def close(self):
        self._flush()
        with open(os.path.join(self.out_dir, "index.json"), "w") as f:
            json.dump({"version": 2, "shards": self.shards}, f)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read_shard(path):
    """Yields one dict per sample of an uncompressed MDS-layout shard."""
    buf = open(path, "rb").read()
    n = int(np.frombuffer(buf, dtype=np.uint32, count=1)[0])
    offsets = np.frombuffer(buf, dtype=np.uint32, count=n + 1, offset=4)
    cfg = json.loads(buf[4 + 4 * (n + 1): int(offsets[0])].decode("utf-8"))
    if cfg.get("compression"):
        raise NotImplementedError(f"compressed shards ({cfg['compression']}) need the codec, which this image does not have")
    names = cfg["column_names"]
    for i in range(n):
        s = memoryview(buf)[int(offsets[i]): int(offsets[i + 1])]
        sizes = np.frombuffer(s, dtype=np.uint32, count=len(names))
        pos, out = 4 * len(names), {}
        for name, size in zip(names, sizes):
            out[name] = _decode_ndarray(bytes(s[pos: pos + int(size)]))
            pos += int(size)
        yield out


class LatentClips(torch.utils.data.IterableDataset):
    """edm2/cs_dataloading.py:53-71 (CsVaeDataset) over a local shard directory: every episode's `mean` [c, t, h, w] is viewed
    as [t, c, h, w] and cut into consecutive `clip_size`-frame clips together with its actions; the tail is dropped.
    `rank` / `world` stride the episodes across data-parallel ranks (StreamingDataset partitions samples the same way)."""

    def __init__(self, local, clip_size, rank=0, world=1):
        self.local, self.clip_size, self.rank, self.world = local, clip_size, rank, world
        with open(os.path.join(local, "index.json")) as f:
            self.index = json.load(f)

    def __iter__(self):
        k = 0
        for shard in self.index["shards"]:
            for ex in read_shard(os.path.join(self.local, shard["raw_data"]["basename"])):
                if k % self.world == self.rank:
                    means = torch.from_numpy(np.array(ex["mean"])).permute(1, 0, 2, 3)
                    actions = torch.from_numpy(np.array(ex["action"]))
                    while means.shape[0] >= self.clip_size:
                        yield means[:self.clip_size], actions[:self.clip_size]
                        means, actions = means[self.clip_size:], actions[self.clip_size:]
                k += 1


def collate(batch):
    """edm2/cs_dataloading.py:77-81 (CsVaeCollate)."""
    means, actions = zip(*batch)
    return torch.stack(means), torch.stack(actions)


def normalize_latents(means, mean, std):
    """cs_train.py:102: latents = (means - vae.mean[:, None, None]) / vae.std[:, None, None], on the device, fp32."""
    return (means.float() - mean[:, None, None]) / std[:, None, None]
