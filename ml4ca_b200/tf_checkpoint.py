"""Reader for TensorFlow-1 "bundle" checkpoints (``variables.index`` + ``variables.data-00000-of-00001``)
without TensorFlow -- enough to load the reference's shipped actor/critic weights.

The reference saves its policies with ``tf.saved_model.simple_save`` (spinup/utils/logx.py:213-229) and
reloads them for inference in spinup/utils/test_policy.py:10-95 and
src/rl/ROS/rl_allocator/src/utils.py:40-86.  The index file is a LevelDB-format table (sorted, prefix
compressed keys, varint-coded block handles) whose values are ``BundleEntryProto`` messages
(dtype, shape, shard, offset, size); the data file is the raw little-endian tensor bytes.
"""
import os
import struct

import numpy as np

_TABLE_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _block_entries(block):
    """Yield (key, value) of one table block (restart array at the end is ignored)."""
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _read_block(data, offset, size):
    # block contents are followed by a 1-byte compression type and a 4-byte crc
    if data[offset + size] != 0:
        raise ValueError("compressed checkpoint index blocks are not supported")
    return data[offset:offset + size]


def _parse_entry(value):
    """BundleEntryProto: 1 dtype, 2 shape (TensorShapeProto: repeated dim{1 size}), 3 shard_id, 4 offset, 5 size."""
    pos, out = 0, {"dtype": 0, "shape": [], "shard": 0, "offset": 0, "size": 0}
    while pos < len(value):
        tag, pos = _varint(value, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _varint(value, pos)
            if field == 1:
                out["dtype"] = v
            elif field == 3:
                out["shard"] = v
            elif field == 4:
                out["offset"] = v
            elif field == 5:
                out["size"] = v
        elif wire == 2:
            ln, pos = _varint(value, pos)
            sub = value[pos:pos + ln]
            pos += ln
            if field == 2:
                sp = 0
                while sp < len(sub):
                    t2, sp = _varint(sub, sp)
                    if t2 & 7 == 2:
                        l2, sp = _varint(sub, sp)
                        dim = sub[sp:sp + l2]
                        sp += l2
                        dp = 0
                        while dp < len(dim):
                            t3, dp = _varint(dim, dp)
                            if t3 & 7 == 0:
                                v3, dp = _varint(dim, dp)
                                if t3 >> 3 == 1:
                                    out["shape"].append(v3)
                            elif t3 & 7 == 2:
                                l3, dp = _varint(dim, dp)
                                dp += l3
                    elif t2 & 7 == 0:
                        _, sp = _varint(sub, sp)
        elif wire == 5:
            pos += 4
        elif wire == 1:
            pos += 8
    return out


def load_bundle(prefix):
    """``prefix`` = '<dir>/variables/variables'.  Returns {variable name: ndarray}."""
    with open(prefix + ".index", "rb") as fh:
        idx = fh.read()
    if struct.unpack_from("<Q", idx, len(idx) - 8)[0] != _TABLE_MAGIC:
        raise ValueError("not a TensorFlow bundle index: %s.index" % prefix)
    footer = idx[-48:]
    pos = 0
    _, pos = _varint(footer, pos)          # metaindex handle
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)       # index handle
    isize, pos = _varint(footer, pos)
    entries = {}
    for _, handle in _block_entries(_read_block(idx, ioff, isize)):
        boff, hp = _varint(handle, 0)
        bsize, hp = _varint(handle, hp)
        for key, value in _block_entries(_read_block(idx, boff, bsize)):
            if key:                        # the empty key holds the BundleHeaderProto
                entries[key.decode()] = _parse_entry(value)
    with open(prefix + ".data-00000-of-00001", "rb") as fh:
        data = fh.read()
    out = {}
    for name, e in entries.items():
        dt = _DTYPES.get(e["dtype"])
        if dt is None or e["shard"] != 0:
            continue
        out[name] = np.frombuffer(data, dtype=dt, count=e["size"] // np.dtype(dt).itemsize,
                                  offset=e["offset"]).reshape(e["shape"]).copy()
    return out


def actor_critic_params(variables):
    """Flatten a loaded bundle into the parameter order of the C ABI (ml4ca_policy_create):
    pi/dense{,_1,..}/{kernel,bias}, pi/log_std, v/dense{,_1,..}/{kernel,bias}
    (variable names as built by core.py:29-33,80-107).  Returns (flat float32, dims dict)."""
    def layers(scope):
        names = sorted({k.rsplit("/", 1)[0] for k in variables if k.startswith(scope + "/dense") and "Adam" not in k},
                       key=lambda s: int(s.split("_")[1]) if "_" in s.split("/")[1] else 0)
        return [(variables[n + "/kernel"], variables[n + "/bias"]) for n in names]
    pi, v = layers("pi"), layers("v")
    log_std = variables["pi/log_std"]
    flat = []
    for W, b in pi:
        flat += [W.ravel(), b.ravel()]
    flat.append(log_std.ravel())
    for W, b in v:
        flat += [W.ravel(), b.ravel()]
    dims = {"obs_dim": int(pi[0][0].shape[0]), "act_dim": int(pi[-1][0].shape[1]), "hidden": int(pi[0][0].shape[1]),
            "n_hidden": len(pi) - 1}
    return np.concatenate(flat).astype(np.float32), dims


def load_actor_critic(save_dir):
    """``save_dir`` = a ``tf1_save`` directory of the reference.  Returns (flat params, dims)."""
    return actor_critic_params(load_bundle(os.path.join(save_dir, "variables", "variables")))
