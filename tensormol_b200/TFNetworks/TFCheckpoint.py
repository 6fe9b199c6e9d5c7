"""Reading the reference's trained networks without TensorFlow.

The reference keeps a trained network in three pieces (TFNetworks/TFManage.py:64-79, TFMolInstance.py:98-124,
TFMolInstanceDirect.py:2212-2215): `<networks>/<manager>.tfm` (pickle of the manager's __dict__, names the instance),
`<networks>/<instance>.tfn` (pickle of the instance's __dict__, holds `chk_file`) and the `tf.train.Saver` checkpoint
`<networks>/<instance>/<instance>-chk-<step>` = a TensorFlow *tensor bundle* (`.index` + `.data-00000-of-00001`).
This module reads all three with numpy only:

* `read_bundle(prefix)` -- the bundle format as published in TensorFlow's sources (TensorFlow itself is a third-party,
  un-vendored dependency of the reference, no version pinned -- README.md:42 "TensorFlow(>1.1)"; the V2 "tensor bundle"
  layout is what `tf.train.Saver` has written by default since TF 0.12, core/util/tensor_bundle): `.index` is an
  SSTable in the LevelDB table format (core/lib/io/format.cc, block.cc: prefix-compressed key/value blocks with restart
  arrays, 5-byte block trailers = compression type + masked CRC-32C, a 48-byte footer ending in the magic
  0xdb4775248b80fb57) whose values are `BundleEntryProto` messages (dtype, shape, shard_id, offset, size, crc32c;
  key "" = `BundleHeaderProto`), and the data shards hold the raw little-endian tensor bytes.
* `write_bundle(prefix, tensors)` -- the inverse (used by the tests, and to hand weights back to a TensorFlow user).
* `weights_from_variables` / `variables_from_weights` -- the reference's variable names for the BP+EE instance
  (TFMolInstanceDirect.py:5164-5272: scopes `EnergyNet/<Z>_hidden<l>`, `DipoleNet/<Z>_hidden<l>_charge`, ...) <-> the
  weight dictionary of this package.
* `load_tm_pickle(path)` -- the `.tfm` / `.tfn` pickles (Python-2 or -3, Containers/PickleTM.py) as plain dictionaries,
  with stand-ins for the reference's classes.

PARITY STATUS: no TensorFlow-written checkpoint exists in this environment (the reference ships no network files and
TensorFlow cannot be installed offline), so the reader is verified against this module's own writer and the published
format only -- "unpinned against a real file"; block and tensor CRCs are checked on read, so a misunderstanding of the
layout shows up as an error rather than as wrong weights.
"""
from __future__ import annotations

import io
import os
import pickle
import re
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_MASK_DELTA = 0xa282ead8
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_, 17: np.uint16,
           19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


class CheckpointError(RuntimeError):
    pass


# ---------------------------------------------------------------------------------------------- CRC-32C (Castagnoli)
def _make_table():
    t = np.zeros(256, np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t[i] = c
    return t


_T0 = _make_table()
_TBL = [int(x) for x in _T0]


def crc32c(data, crc=0):
    c = crc ^ 0xFFFFFFFF
    tbl = _TBL
    for b in bytes(data):
        c = tbl[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data):
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- varints / protobuf wire
def _get_varint(buf, pos):
    shift = result = 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _parse_message(buf):
    """Flat protobuf wire decode -> {field: [values]} (varint ints, fixed32 ints, length-delimited bytes)."""
    out, pos = {}, 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError(f"unsupported protobuf wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_entry(buf):
    m = _parse_message(buf)
    shape = []
    for sh in m.get(2, []):                                  # TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}
        for d in _parse_message(sh).get(2, []):
            shape.append(_signed64(_parse_message(d).get(1, [0])[0]))
    if 7 in m:
        raise CheckpointError("sliced (partitioned) variables are not supported")
    return dict(dtype=m.get(1, [0])[0], shape=tuple(shape), shard_id=m.get(3, [0])[0], offset=m.get(4, [0])[0],
                size=m.get(5, [0])[0], crc32c=m.get(6, [None])[0])


def _field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


def _encode_entry(dtype_id, shape, shard_id, offset, size, crc):
    dims = b"".join(_field(2, 2, (lambda d: _put_varint(len(d)) + d)(_field(1, 0, _put_varint(int(s))))) for s in shape)
    out = _field(1, 0, _put_varint(dtype_id)) + _field(2, 2, _put_varint(len(dims)) + dims)
    if shard_id:
        out += _field(3, 0, _put_varint(shard_id))
    if offset:
        out += _field(4, 0, _put_varint(offset))
    out += _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return out


# ---------------------------------------------------------------------------------------------- SSTable (LevelDB table format)
def _read_block(buf, offset, size, verify=True):
    if offset + size + 5 > len(buf):
        raise CheckpointError("block handle outside the file")
    contents = buf[offset:offset + size]
    ctype = buf[offset + size]
    stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
    if verify and stored != masked_crc32c(buf[offset:offset + size + 1]):
        raise CheckpointError("block checksum mismatch in the .index file")
    if ctype != 0:
        raise CheckpointError("compressed index blocks (snappy) are not supported; tf.train.Saver writes them uncompressed")
    return contents


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointError("bad block")
    nrestart = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestart
    if end < 0:
        raise CheckpointError("bad restart array")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key):
            raise CheckpointError("bad key prefix")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _build_block(items, restart_interval=16):
    out, restarts, last = bytearray(), [], b""
    for n, (k, v) in enumerate(items):
        if n % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def read_index(path, verify=True):
    """`.index` file -> (header dict, {tensor name: entry dict})."""
    with open(path, "rb") as fh:
        buf = fh.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise CheckpointError(f"{path}: not a TensorFlow V2 checkpoint index (table magic missing); V1 checkpoints are not supported")
    foot = buf[len(buf) - 48:len(buf) - 8]
    _mo, p = _get_varint(foot, 0)
    _ms, p = _get_varint(foot, p)
    io_, p = _get_varint(foot, p)
    is_, p = _get_varint(foot, p)
    entries, header = {}, None
    for _k, handle in _block_entries(_read_block(buf, io_, is_, verify)):
        bo, q = _get_varint(handle, 0)
        bs, q = _get_varint(handle, q)
        for k, v in _block_entries(_read_block(buf, bo, bs, verify)):
            if k == b"":
                m = _parse_message(v)
                header = dict(num_shards=m.get(1, [1])[0], endianness=m.get(2, [0])[0])
            else:
                entries[k.decode("utf-8")] = _parse_entry(v)
    if header is None:
        raise CheckpointError(f"{path}: bundle header entry missing")
    if header["endianness"] != 0:
        raise CheckpointError("big-endian bundles are not supported")
    return header, entries


def read_bundle(prefix, names=None, verify="auto"):
    """All (or the named) tensors of the checkpoint `<prefix>.index` / `<prefix>.data-*` as numpy arrays.
    verify: True = check every tensor's CRC-32C, "auto" = tensors up to 256 KiB (pure-Python CRC), False = none."""
    header, entries = read_index(prefix + ".index", verify is not False)
    shards, out = {}, {}
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e["dtype"] not in _DTYPES:
            continue                                          # strings, resources ...: not weights
        sid = e["shard_id"]
        if sid not in shards:
            p = "%s.data-%05d-of-%05d" % (prefix, sid, header["num_shards"])
            if not os.path.exists(p):
                raise CheckpointError(f"missing data shard {p}")
            shards[sid] = np.memmap(p, np.uint8, "r") if os.path.getsize(p) else np.zeros(0, np.uint8)
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        dt = np.dtype(_DTYPES[e["dtype"]])
        n = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if len(raw) != e["size"] or n * dt.itemsize != e["size"]:
            raise CheckpointError(f"{name}: {e['size']} bytes for shape {e['shape']} {dt}")
        if e["crc32c"] is not None and (verify is True or (verify == "auto" and e["size"] <= (1 << 18))):
            if masked_crc32c(raw.tobytes()) != e["crc32c"]:
                raise CheckpointError(f"{name}: tensor checksum mismatch")
        out[name] = np.frombuffer(raw.tobytes(), dt).reshape(e["shape"]).copy()
    if names is not None:
        missing = [n for n in names if n not in out]
        if missing:
            raise CheckpointError("variables not in the checkpoint: " + ", ".join(missing))
    return out


def write_bundle(prefix, tensors):
    """{name: ndarray} -> `<prefix>.index` + `<prefix>.data-00000-of-00001` in the layout tf.train.Saver writes
    (one shard, keys sorted, uncompressed 4 KiB blocks)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = [(b"", _field(1, 0, _put_varint(1)) + _field(3, 2, (lambda d: _put_varint(len(d)) + d)(_field(1, 0, _put_varint(1)))))]
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            a = np.asarray(tensors[name])
            a = a if a.flags.c_contiguous else np.ascontiguousarray(a)     # (ascontiguousarray would make a scalar 1-d)
            if a.dtype not in _DTYPE_IDS:
                raise CheckpointError(f"{name}: dtype {a.dtype} not supported")
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            fh.write(raw)
            items.append((name.encode("utf-8"), _encode_entry(_DTYPE_IDS[a.dtype], a.shape, 0, offset, len(raw), masked_crc32c(raw))))
            offset += len(raw)
    out = bytearray()

    def emit(block):
        handle = _put_varint(len(out)) + _put_varint(len(block))
        out.extend(block)
        out.append(0)
        out.extend(struct.pack("<I", masked_crc32c(block + b"\x00")))
        return handle

    index_items, cur, cur_size = [], [], 0
    for k, v in items:
        cur.append((k, v))
        cur_size += len(k) + len(v) + 3
        if cur_size >= 4096:
            index_items.append((cur[-1][0], emit(_build_block(cur))))
            cur, cur_size = [], 0
    if cur:
        index_items.append((cur[-1][0], emit(_build_block(cur))))
    meta = emit(_build_block([]))
    index = emit(_build_block(index_items, restart_interval=1))
    foot = meta + index
    out.extend(foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", TABLE_MAGIC))
    with open(prefix + ".index", "wb") as fh:
        fh.write(bytes(out))


# ---------------------------------------------------------------------------------------------- variable names <-> weights
def variable_names(eles, n_hidden):
    """The reference graph's variable names for the BP+EE instance (TFMolInstanceDirect.py:5164-5272), as
    {(net, Z, layer): (weights name, biases name)}; layer == n_hidden is the linear output layer."""
    out = {}
    for z in eles:
        for l in range(n_hidden):
            out[("energy", int(z), l)] = (f"EnergyNet/{z}_hidden{l + 1}/weights", f"EnergyNet/{z}_hidden{l + 1}/biaseslayer{l}")
            out[("charge", int(z), l)] = (f"DipoleNet/{z}_hidden{l + 1}_charge/weights", f"DipoleNet/{z}_hidden{l + 1}_charge/biases")
        out[("energy", int(z), n_hidden)] = (f"EnergyNet/{z}_regression_linear/weights", f"EnergyNet/{z}_regression_linear/biases")
        out[("charge", int(z), n_hidden)] = (f"DipoleNet/{z}_regression_linear_charge/weights", f"DipoleNet/{z}_regression_linear_charge/biases")
    return out


def _find(variables, name):
    """Exact name, else the unique variable whose name ends with it after an outer scope (e.g. a tower prefix);
    optimiser slots (`.../Adam`, `.../Adam_1`) never match because the leaf name differs."""
    if name in variables:
        return variables[name]
    hits = [k for k in variables if k.endswith("/" + name)]
    if len(hits) == 1:
        return variables[hits[0]]
    raise CheckpointError(("variable not found: " if not hits else "ambiguous variable: ") + name)


def weights_from_variables(variables, eles, hidden, inshape=None):
    """{variable name: array} -> {"charge": {Z: [(W, b), ...]}, "energy": {...}} (float64), shapes checked."""
    nh = len(hidden)
    names = variable_names(eles, nh)
    out = {"charge": {}, "energy": {}}
    for net in ("charge", "energy"):
        for z in eles:
            layers = []
            for l in range(nh + 1):
                wn, bn = names[(net, int(z), l)]
                W = np.asarray(_find(variables, wn), np.float64)
                b = np.asarray(_find(variables, bn), np.float64).reshape(-1)
                rows = (inshape if l == 0 else hidden[l - 1])
                cols = hidden[l] if l < nh else 1
                if W.ndim != 2 or W.shape[1] != cols or (rows is not None and W.shape[0] != rows) or b.shape[0] != cols:
                    raise CheckpointError(f"{wn}: shape {W.shape} / bias {b.shape}, expected [{rows}, {cols}]")
                layers.append((W, b))
            out[net][int(z)] = layers
    return out


def variables_from_weights(weights, dtype=np.float64):
    eles = sorted(weights["energy"])
    nh = len(weights["energy"][eles[0]]) - 1
    names = variable_names(eles, nh)
    out = {}
    for net in ("charge", "energy"):
        for z in eles:
            for l, (W, b) in enumerate(weights[net][z]):
                wn, bn = names[(net, z, l)]
                out[wn] = np.asarray(W, dtype)
                out[bn] = np.asarray(b, dtype).reshape(-1)
    return out


def latest_checkpoint(train_dir, name=None):
    """`<train_dir>/<name>-chk-<largest step>` (TFInstance.py:252-259 looks for the .meta files; here the .index files)."""
    best = None
    for f in os.listdir(train_dir):
        m = re.match(r"(.*-chk-(\d+))\.index$", f)
        if m and (name is None or m.group(1).startswith(name + "-chk-")):
            if best is None or int(m.group(2)) > best[0]:
                best = (int(m.group(2)), os.path.join(train_dir, m.group(1)))
    return best[1] if best else None


# ---------------------------------------------------------------------------------------------- .tfm / .tfn pickles
class _Stub:
    """Stand-in for any class of the reference (or of a library that is not installed) found in a pickle."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        elif isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):
            self.__dict__.update(state[0] or {})
            self.__dict__.update(state[1])
        else:
            self.__dict__["_state"] = state


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("numpy", "builtins", "__builtin__", "copy_reg", "copyreg", "collections", "_codecs"):
            return super().find_class(module, name)
        return type(name, (_Stub,), {"__module__": module})


def load_tm_pickle(path):
    """A `.tfm` / `.tfn` file as a plain dict (Containers/PickleTM.py:44-60 reads Python-2 pickles with latin-1 strings)."""
    with open(path, "rb") as fh:
        data = fh.read()
    obj = _TolerantUnpickler(io.BytesIO(data), encoding="latin1").load()
    if not isinstance(obj, dict):
        obj = dict(getattr(obj, "__dict__", {}))
    return obj


def find_reference_network(name, networks_dir):
    """Manager name -> checkpoint prefix, following TFMolManage.Prepare / MolInstance.Load (TFMolManage.py:1468-1470,
    TFMolInstance.py:104-113): `<dir>/<name>.tfm` names the instance, `<dir>/<instance>.tfn` holds chk_file (its
    "./networks/" prefix is re-rooted on networks_dir); falls back to the newest checkpoint in `<dir>/<instance>/`.
    Returns (prefix, instance dict) or (None, None)."""
    tfm = os.path.join(networks_dir, name + ".tfm")
    if not os.path.exists(tfm):
        return None, None
    mgr = load_tm_pickle(tfm)
    trained = mgr.get("TrainedNetworks") or []
    if not trained:
        return None, None
    inst_name = trained[0]
    inst = {}
    tfn = os.path.join(networks_dir, inst_name + ".tfn")
    if os.path.exists(tfn):
        inst = load_tm_pickle(tfn)
    chk = inst.get("chk_file")
    if isinstance(chk, bytes):
        chk = chk.decode("latin1")
    if chk:
        chk = chk.replace("./networks/", networks_dir if networks_dir.endswith("/") else networks_dir + "/")
        if not os.path.exists(chk + ".index"):
            alt = os.path.join(networks_dir, inst_name, os.path.basename(chk))
            chk = alt if os.path.exists(alt + ".index") else None
    if not chk and os.path.isdir(os.path.join(networks_dir, inst_name)):
        chk = latest_checkpoint(os.path.join(networks_dir, inst_name))
    return chk, inst
