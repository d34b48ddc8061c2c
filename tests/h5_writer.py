"""Minimal HDF5 writer for the reader's tests (TEST INFRASTRUCTURE ONLY).

h5py does not exist in this image, so the fixtures of ``tests/test_hdf5_reader.py`` are produced here, straight from the HDF5
File Format Specification (version 3.0) and independently of the reader's code paths: superblock version 0, a root group as a
symbol table (local heap + version-1 B-tree + one symbol-table node), version-1 object headers, contiguous / chunked
(optionally deflate + shuffle) datasets, attributes with scalar float / int, fixed-length and variable-length strings (global
heap) -- the structures ``h5py.File(path, "w")`` + ``create_dataset(name, data=array)`` + ``f.attrs[k] = v`` create with
libhdf5's default ``libver="earliest"`` (``src/diffusion_pde/pdes/utils.py:112-127`` of the reference).
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    order = 1 if dt.byteorder == ">" else 0
    if dt.kind == "f":
        size = dt.itemsize
        exp_bits, mant_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[size]
        bits = order | 0x20 | ((8 * size - 1) << 8)                  # mantissa normalisation "implied", sign bit location
        props = struct.pack("<HHBBBBI", 0, 8 * size, mant_bits, exp_bits, 0, mant_bits, bias)
        return struct.pack("<B", 0x11) + bits.to_bytes(3, "little") + struct.pack("<I", size) + props
    if dt.kind in "iu":
        bits = order | (8 if dt.kind == "i" else 0)
        return struct.pack("<B", 0x10) + bits.to_bytes(3, "little") + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise TypeError(dt)


def _str_dtype(n: int) -> bytes:
    return struct.pack("<B", 0x13) + (0).to_bytes(3, "little") + struct.pack("<I", n)


def _vlen_str_dtype() -> bytes:
    return struct.pack("<B", 0x19) + (0x01 | (1 << 8)).to_bytes(3, "little") + struct.pack("<I", 16) + _str_dtype(1)


def _dataspace(shape) -> bytes:
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _message(mtype: int, body: bytes) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _object_header(messages) -> bytes:
    blob = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(blob)) + blob


class Writer:
    def __init__(self):
        self.buf = bytearray(96)                                     # superblock written last

    def _alloc(self, data: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    @staticmethod
    def _vlen_payloads(attr_dicts):
        return [v[5:].encode() for d in attr_dicts for v in d.values() if isinstance(v, str) and v.startswith("vlen:")]

    def _attr_messages(self, attrs, gheap_addr, counter):
        """Attribute messages (version 1); `counter[0]` = index of the next variable-length string in the global heap."""
        out = []
        for name, value in attrs.items():
            nm = name.encode() + b"\0"
            if isinstance(value, str) and value.startswith("vlen:"):
                dt, sp = _vlen_str_dtype(), _dataspace(())
                data = struct.pack("<IQI", len(value[5:].encode()), gheap_addr, counter[0])
                counter[0] += 1
            elif isinstance(value, str):
                raw = value.encode() + b"\0"
                dt, sp, data = _str_dtype(len(raw)), _dataspace(()), raw
            else:
                arr = np.asarray(value)
                dt, sp, data = _dtype_msg(arr.dtype), _dataspace(arr.shape), arr.tobytes()
            head = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp)
            out.append(_message(0x0C, head + data))
        return out

    def _dataset_messages(self, arr: np.ndarray, chunks=None, deflate=False, shuffle=False):
        arr = np.ascontiguousarray(arr)
        msgs = [_message(0x01, _dataspace(arr.shape)), _message(0x03, _dtype_msg(arr.dtype))]
        if chunks is None:
            addr = self._alloc(arr.tobytes())
            msgs.append(_message(0x08, struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)))
            return msgs
        filters = ([(2, [arr.dtype.itemsize])] if shuffle else []) + ([(1, [4])] if deflate else [])
        entries = []
        grid = [range(0, s, c) for s, c in zip(arr.shape, chunks)]
        for offs in np.stack(np.meshgrid(*grid, indexing="ij"), -1).reshape(-1, arr.ndim):
            block = np.zeros(chunks, arr.dtype)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, arr.shape))
            block[tuple(slice(0, s.stop - s.start) for s in sel)] = arr[sel]
            raw = block.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, np.uint8).reshape(-1, arr.dtype.itemsize).T.tobytes()
            if deflate:
                raw = zlib.compress(raw, 4)
            entries.append((list(offs), len(raw), self._alloc(raw)))
        node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF)
        for offs, size, caddr in entries:                            # key (size, filter mask, offsets + 0), child address
            node += struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", int(o)) for o in offs) + struct.pack("<Q", 0)
            node += struct.pack("<Q", caddr)
        node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", int(s)) for s in arr.shape) + struct.pack("<Q", 0)
        bt = self._alloc(node)
        if filters:
            fm = struct.pack("<BB6x", 1, len(filters))
            for fid, vals in filters:
                fm += struct.pack("<HHHH", fid, 0, 1, len(vals)) + b"".join(struct.pack("<I", v) for v in vals)
                if len(vals) % 2:
                    fm += b"\0" * 4
            msgs.append(_message(0x0B, fm))
        msgs.append(_message(0x08, struct.pack("<BBBQ", 3, 2, arr.ndim + 1, bt) +
                             b"".join(struct.pack("<I", c) for c in list(chunks) + [arr.dtype.itemsize])))
        return msgs

    def write(self, path, datasets: dict, attrs: dict, dataset_kw: dict | None = None, dataset_attrs: dict | None = None):
        dataset_kw, dataset_attrs = dataset_kw or {}, dataset_attrs or {}
        names = sorted(datasets)
        # global heap collection holding every variable-length string, in the order the attributes are encoded below
        payloads = self._vlen_payloads([attrs] + [dataset_attrs.get(n, {}) for n in names])
        gheap_addr = 0
        if payloads:
            body = b"".join(struct.pack("<HH4xQ", i, 1, len(s)) + _pad8(s) for i, s in enumerate(payloads, 1))
            size = 16 + len(body) + 16
            gheap_addr = self._alloc(b"GCOL" + struct.pack("<B3xQ", 1, size) + body + struct.pack("<HH4xQ", 0, 0, 0))
        counter = [1]
        root_attr_msgs = self._attr_messages(attrs, gheap_addr, counter)
        addr = {}
        for n in names:
            msgs = self._dataset_messages(np.asarray(datasets[n]), **dataset_kw.get(n, {}))
            msgs += self._attr_messages(dataset_attrs.get(n, {}), gheap_addr, counter)
            addr[n] = self._alloc(_object_header(msgs))
        # local heap with the link names, one symbol-table node, one B-tree node
        heap_data, offs = b"\0" * 8, {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += _pad8(n.encode() + b"\0")
        heap_data_addr = self._alloc(heap_data)
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, heap_data_addr))
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for n in names:
            snod += struct.pack("<QQII16x", offs[n], addr[n], 0, 0)
        snod_addr = self._alloc(snod)
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, offs[names[-1]] if names else 0)
        tree_addr = self._alloc(tree)
        root = self._alloc(_object_header([_message(0x11, struct.pack("<QQ", tree_addr, heap_addr))] + root_attr_msgs))
        eof = len(self.buf) + (-len(self.buf) % 8)
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree_addr, heap_addr)
        assert len(sb) == 96
        self.buf[:96] = sb
        with open(path, "wb") as fh:
            fh.write(bytes(self.buf) + b"\0" * (eof - len(self.buf)))


def write_h5(path, datasets, attrs=None, dataset_kw=None, dataset_attrs=None):
    Writer().write(path, datasets, attrs or {}, dataset_kw, dataset_attrs)
