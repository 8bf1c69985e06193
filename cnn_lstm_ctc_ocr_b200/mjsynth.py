"""The training-side input format of the reference (src/weinman/mjsynth.py), without TensorFlow.

  read_tfrecord(path)              <- tf.TFRecordReader                      (mjsynth.py:159-160)
  parse_example(serialized)        <- tf.parse_single_example                (mjsynth.py:162-176)
  read_word_record(serialized)     <- mjsynth._read_word_record              (mjsynth.py:157-183)
  preprocess_image(image)          <- mjsynth._preprocess_image              (mjsynth.py:185-194)  [host form]
  bucketed_input_pipeline(...)     <- mjsynth.bucketed_input_pipeline        (mjsynth.py:28-73)
  threaded_input_pipeline(...)     <- mjsynth.threaded_input_pipeline        (mjsynth.py:75-107)
  write_tfrecord / make_example    <- mjsynth-tfrecord.py:154-175            (fixtures and round-trip tests)

Record framing (TFRecord): uint64 length | uint32 masked CRC32C(length) | payload | uint32 masked CRC32C(payload), little
endian.  Payload: a protobuf `tf.train.Example{ features{ map<string, Feature> } }`, Feature = oneof{bytes_list=1,
float_list=2, int64_list=3}; parsed here with a 40-line varint reader (no protobuf runtime, no TensorFlow).

The pipelines are host code by nature (file I/O, JPEG decode, Python queues in the reference).  What reaches the GPU is a
batch in the reference's own contract: image float32 [B,32,Wmax,1] preprocessed and padded with 0.0 to the widest crop of
the batch (dynamic_pad pads the PREPROCESSED tensor: mjsynth.py:56,69 -- unlike the serving side, which pads uint8 0 =
-0.5, server.py:30), width int32 [B], label SparseTensor (int32 values), length, text, filename.  `as_uint8=True` hands out
the raw uint8 rows + widths instead, for Trainer's device-side preprocessing (`ocr_preprocess_train`: same arithmetic
fused into one kernel; tests assert both forms agree bit for bit).

TensorFlow's queue runners make the reference's batch ORDER nondeterministic (4 reader threads, shuffled file names); the
stand-in is deterministic: records in file order (optionally a seeded shuffle of the file list per epoch), one batch
leaves a bucket the moment it holds batch_size records -- the order a single-threaded TF pipeline would produce.
"""
import glob
import io
import os
import struct

import numpy as np

from .ctc import SparseTensor

# ----------------------------------------------------------------------------- CRC32C (Castagnoli), masked as TFRecord does
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.zeros(256, np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t[i] = c
        _CRC_TABLE = t.tolist()
    return _CRC_TABLE


def crc32c(data):
    t = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data):
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- TFRecord framing
def read_tfrecord(path, verify=False):
    """Yield the serialized payload of every record.  verify=True checks both CRCs (pure Python: slow, tests only)."""
    with open(path, "rb") as f:
        buf = f.read()
    pos, n = 0, len(buf)
    while pos < n:
        if pos + 12 > n:
            raise ValueError("%s: truncated record header at byte %d" % (path, pos))
        (length,) = struct.unpack_from("<Q", buf, pos)
        (lcrc,) = struct.unpack_from("<I", buf, pos + 8)
        if pos + 12 + length + 4 > n:
            raise ValueError("%s: truncated record at byte %d" % (path, pos))
        payload = buf[pos + 12:pos + 12 + length]
        if verify:
            (dcrc,) = struct.unpack_from("<I", buf, pos + 12 + length)
            if masked_crc32c(buf[pos:pos + 8]) != lcrc or masked_crc32c(payload) != dcrc:
                raise ValueError("%s: corrupted record at byte %d" % (path, pos))
        yield payload
        pos += 12 + length + 4


def write_tfrecord(path, payloads):
    with open(path, "wb") as f:
        for p in payloads:
            hdr = struct.pack("<Q", len(p))
            f.write(hdr + struct.pack("<I", masked_crc32c(hdr)) + p + struct.pack("<I", masked_crc32c(p)))


# ----------------------------------------------------------------------------- protobuf: tf.train.Example
def _varint(b, i):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        s += 7
        if c < 0x80:
            return r, i


def _fields(b):
    """(field number, wire type, value) of one message; value = int (varint), bytes (length-delimited / fixed)."""
    i, n = 0, len(b)
    while i < n:
        key, i = _varint(b, i)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(b, i)
        elif wt == 2:
            ln, i = _varint(b, i)
            v = b[i:i + ln]
            i += ln
        elif wt == 5:
            v = b[i:i + 4]
            i += 4
        elif wt == 1:
            v = b[i:i + 8]
            i += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield fn, wt, v


def _int64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def parse_example(serialized):
    """tf.train.Example -> {feature name (str): list of bytes | float | int}."""
    out = {}
    for fn, wt, features in _fields(serialized):
        if fn != 1:
            continue
        for fn2, wt2, entry in _fields(features):          # map<string, Feature> entries
            if fn2 != 1:
                continue
            name, feat = None, b""
            for fn3, wt3, v in _fields(entry):
                if fn3 == 1:
                    name = v.decode("utf-8")
                elif fn3 == 2:
                    feat = v
            vals = []
            for kind, wtk, lst in _fields(feat):
                for fn4, wt4, v in _fields(lst):
                    if fn4 != 1:
                        continue
                    if kind == 1:                             # BytesList
                        vals.append(bytes(v))
                    elif kind == 2:                           # FloatList: packed or repeated fixed32
                        vals.extend(struct.unpack("<%df" % (len(v) // 4), v))
                    elif kind == 3:                           # Int64List: packed varints or one varint per entry
                        if wt4 == 2:
                            j = 0
                            while j < len(v):
                                x, j = _varint(v, j)
                                vals.append(_int64(x))
                        else:
                            vals.append(_int64(v))
            out[name] = vals
    return out


def _enc_varint(x):
    x &= (1 << 64) - 1
    out = bytearray()
    while True:
        if x < 0x80:
            out.append(x)
            return bytes(out)
        out.append((x & 0x7F) | 0x80)
        x >>= 7


def _ld(fn, payload):
    return _enc_varint((fn << 3) | 2) + _enc_varint(len(payload)) + payload


def make_example(features):
    """{name: list of bytes | list of int} -> serialized tf.train.Example (mjsynth-tfrecord.py:166-174 writes
    image/encoded, image/labels, image/height, image/width, image/filename, text/string, text/length)."""
    entries = b""
    for name in sorted(features):
        vals = features[name]
        if vals and isinstance(vals[0], (bytes, bytearray)):
            feat = _ld(1, b"".join(_ld(1, bytes(v)) for v in vals))
        elif vals and isinstance(vals[0], float):
            feat = _ld(2, _ld(1, struct.pack("<%df" % len(vals), *vals)))
        else:
            feat = _ld(3, _ld(1, b"".join(_enc_varint(int(v)) for v in vals)))
        entries += _ld(1, _ld(1, name.encode("utf-8")) + _ld(2, feat))
    return _ld(1, entries)


# ----------------------------------------------------------------------------- mjsynth._read_word_record
def decode_jpeg_gray(encoded):
    """tf.image.decode_jpeg(channels=1): libjpeg's own grayscale output (the luma plane), uint8 [H,W,1].  PIL drives the
    same libjpeg; its IDCT/upsampling defaults may differ from TensorFlow's by +-1 LSB (not a parity surface)."""
    from PIL import Image
    im = Image.open(io.BytesIO(encoded))
    im.draft("L", im.size)
    return np.asarray(im.convert("L"), dtype=np.uint8)[:, :, None]


def read_word_record(serialized):
    """-> (image uint8 [H,W,1], width int32 [1], labels int64 [L], length int64 [1], text bytes, filename bytes) with the
    feature_map defaults of mjsynth.py:162-175 (missing strings '', missing width / length 1)."""
    f = parse_example(serialized)
    enc = f.get("image/encoded", [b""])[0]
    image = decode_jpeg_gray(enc)
    width = np.asarray(f.get("image/width", [1])[:1], np.int32)
    labels = np.asarray(f.get("image/labels", []), np.int64)
    length = np.asarray(f.get("text/length", [1])[:1], np.int64)
    text = f.get("text/string", [b""])[0]
    filename = f.get("image/filename", [b""])[0]
    return image, width, labels, length, text, filename


def preprocess_image(image):
    """mjsynth._preprocess_image (mjsynth.py:185-194) on the host: convert_image_dtype (uint8 * float32(1/255)), - 0.5,
    then a copy of the first row on top (31 -> 32 rows).  image uint8 [H,W,1] -> float32 [H+1,W,1]."""
    x = image.astype(np.float32) * np.float32(1.0 / 255.0) - np.float32(0.5)
    return np.concatenate([x[:1], x], 0)


def _keep(width, width_threshold, length, length_threshold):
    """mjsynth._get_input_filter (mjsynth.py:109-141)."""
    ok = True
    if width_threshold is not None:
        ok = ok and int(width) <= width_threshold
    if length_threshold is not None:
        ok = ok and int(length) <= length_threshold
    return ok


def _data_files(base_dir, file_patterns):
    files = [f for pat in file_patterns for f in sorted(glob.glob(os.path.join(base_dir, pat)))]
    if not files:
        raise ValueError("no record files match %s in %s" % (list(file_patterns), base_dir))
    return files


def _records(base_dir, file_patterns, num_epochs, shuffle_seed):
    files = _data_files(base_dir, file_patterns)
    rng = np.random.default_rng(shuffle_seed) if shuffle_seed is not None else None
    epoch = 0
    while num_epochs is None or epoch < num_epochs:
        order = list(files)
        if rng is not None:
            rng.shuffle(order)        # tf.train.string_input_producer shuffles the file names every epoch
        for path in order:
            for payload in read_tfrecord(path):
                yield payload
        epoch += 1


def _make_batch(items, as_uint8):
    """tf.train.batch / bucket_by_sequence_length with dynamic_pad=True: pad every tensor with zeros to the batch maximum;
    the serialized sparse labels become one SparseTensor (tf.deserialize_many_sparse) cast to int32 (mjsynth.py:71-72)."""
    B = len(items)
    widths = np.concatenate([it[1] for it in items]).astype(np.int32)
    if as_uint8:
        H = items[0][0].shape[0]
        wmax = max(it[0].shape[1] for it in items)
        image = np.zeros((B, H, wmax, 1), np.uint8)
        for b, it in enumerate(items):
            image[b, :, :it[0].shape[1]] = it[0]
    else:
        pre = [preprocess_image(it[0]) for it in items]
        H = pre[0].shape[0]
        wmax = max(p.shape[1] for p in pre)
        image = np.zeros((B, H, wmax, 1), np.float32)            # dynamic_pad: 0.0, AFTER preprocessing
        for b, p in enumerate(pre):
            image[b, :, :p.shape[1]] = p
    lens = [len(it[2]) for it in items]
    import torch
    idx = np.array([(b, j) for b, n in enumerate(lens) for j in range(n)], np.int64).reshape(-1, 2)
    vals = np.concatenate([it[2] for it in items]).astype(np.int32) if sum(lens) else np.zeros(0, np.int32)
    label = SparseTensor(torch.from_numpy(idx), torch.from_numpy(vals), torch.tensor([B, max(lens) if lens else 0], dtype=torch.int64))
    length = np.stack([it[3] for it in items]).astype(np.int64)              # [B,1]
    text = [it[4] for it in items]
    filename = [it[5] for it in items]
    return image, widths, label, length, text, filename


def bucketed_input_pipeline(base_dir, file_patterns=("*.tfrecord",), num_threads=4, batch_size=32,
                            boundaries=(32, 64, 96, 128, 160, 192, 224, 256), input_device=None, width_threshold=None,
                            length_threshold=None, num_epochs=None, shuffle_seed=None, as_uint8=False):
    """mjsynth.bucketed_input_pipeline (mjsynth.py:28-73) as a generator of (image, width, label, length, text, filename).

    tf.contrib.training.bucket_by_sequence_length: a record of width w goes to bucket
    #{boundaries <= w} (bucket i holds boundaries[i-1] <= w < boundaries[i]; open-ended first and last buckets); a bucket
    releases a batch when it holds batch_size records; with num_epochs set (allow_smaller_final_batch) what is left at the
    end leaves as smaller batches, in bucket order.  Records failing the width / length thresholds are dropped
    (keep_input).  Images are padded to the widest crop of THEIR batch.  num_threads / input_device are accepted for
    signature compatibility (the queues they configured do not exist here)."""
    bounds = list(boundaries)
    buckets = [[] for _ in range(len(bounds) + 1)]
    for payload in _records(base_dir, file_patterns, num_epochs, shuffle_seed):
        rec = read_word_record(payload)
        if not _keep(rec[1][0], width_threshold, rec[3][0], length_threshold):
            continue
        k = int(np.searchsorted(bounds, int(rec[1][0]), side="right"))
        buckets[k].append(rec)
        if len(buckets[k]) == batch_size:
            yield _make_batch(buckets[k], as_uint8)
            buckets[k] = []
    for bk in buckets:
        if bk:
            yield _make_batch(bk, as_uint8)


def threaded_input_pipeline(base_dir, file_patterns=("*.tfrecord",), num_threads=4, batch_size=32, batch_device=None,
                            preprocess_device=None, num_epochs=None, shuffle_seed=None, as_uint8=False):
    """mjsynth.threaded_input_pipeline (mjsynth.py:75-107): batches in record order, no width bucketing (test.py:66-73
    evaluates with it)."""
    cur = []
    for payload in _records(base_dir, file_patterns, num_epochs, shuffle_seed):
        cur.append(read_word_record(payload))
        if len(cur) == batch_size:
            yield _make_batch(cur, as_uint8)
            cur = []
    if cur:
        yield _make_batch(cur, as_uint8)
