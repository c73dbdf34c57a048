"""TEST INFRASTRUCTURE: writers for minimal DSF and DSDIFF files (the two containers dsf2flac
reads), used to put the same synthetic stream in front of the reference's own file readers
(oracle/_ref) and of the product's container path.

Layouts (reference parsers: dsf_file_reader.cpp:163-288, dsdiff_file_reader.cpp:288-383,1021-1039):
  DSF    "DSD "(28 B: size, file size, metadata pointer; u64 LE) . "fmt "(52 B) . "data"(12 B + blocks of
         block_size bytes per channel, channel after channel, last block zero padded); LSB-first bytes
  DSDIFF "FRM8"(size u64 BE)"DSD " . "FVER" . "PROP"("SND " . "FS  " . "CHNL" . "CMPR") . "DSD "(byte
         interleaved, MSB-first bytes) [. trailing chunk]; chunk sizes exclude the 12-byte header
"""
import struct

import numpy as np


def write_dsf(path, planar, dsd_fs, sample_count=None, block=4096, trailer=b""):
    planar = np.ascontiguousarray(planar, dtype=np.uint8)
    nch, n = planar.shape
    if sample_count is None:
        sample_count = 8 * n
    nb = -(-n // block)
    padded = np.zeros((nch, nb * block), dtype=np.uint8)
    padded[:, :n] = planar
    payload = np.ascontiguousarray(padded.reshape(nch, nb, block).transpose(1, 0, 2)).tobytes()
    data_size = 12 + len(payload)
    total = 28 + 52 + data_size + len(trailer)
    chan_type = {1: 1, 2: 2, 3: 3, 4: 5, 5: 6, 6: 7}.get(nch, 2)
    with open(path, "wb") as f:
        f.write(b"DSD " + struct.pack("<QQQ", 28, total, 0))
        f.write(b"fmt " + struct.pack("<QIIIIIIQII", 52, 1, 0, chan_type, nch, dsd_fs, 1, sample_count, block, 0))
        f.write(b"data" + struct.pack("<Q", data_size))
        f.write(payload)
        f.write(trailer)


def _chunk(ident, body):
    return ident + struct.pack(">Q", len(body)) + body + (b"\0" if len(body) & 1 else b"")


def write_dsdiff(path, planar, dsd_fs, trailer_chunk=False):
    planar = np.ascontiguousarray(planar, dtype=np.uint8)
    nch, n = planar.shape
    ids = [b"SLFT", b"SRGT", b"C   ", b"LFE ", b"LS  ", b"RS  ", b"C007", b"C008"][:nch]
    prop = b"SND " + _chunk(b"FS  ", struct.pack(">I", dsd_fs)) + _chunk(b"CHNL", struct.pack(">H", nch) + b"".join(ids)) \
        + _chunk(b"CMPR", b"DSD " + bytes([14]) + b"not compressed")
    body = b"DSD " + _chunk(b"FVER", struct.pack(">I", 0x01050000)) + _chunk(b"PROP", prop) \
        + _chunk(b"DSD ", np.ascontiguousarray(planar.T).tobytes())
    if trailer_chunk:
        body += _chunk(b"COMT", struct.pack(">H", 0))
    with open(path, "wb") as f:
        f.write(b"FRM8" + struct.pack(">Q", len(body)) + body)
