"""Read-only HDF5 reader for ARTIST scenario files - the subset ``h5py.File`` is used for by the reference's loaders
(``artist/scenario/scenario.py:105-259``, ``artist/io/h5_scenario_parser.py``, ``artist/field/*.from_hdf5``):
``f[key]`` (also ``"a/b/c"`` paths), ``.keys()``, ``len()``, ``in``, ``.get()``, ``dataset[()]``, ``.attrs[...]``.

Neither the build image nor the GPU image has ``h5py``/libhdf5, so the format is parsed directly from the file bytes
(HDF5 File Format Specification v3, the parts scenario files written by h5py's defaults contain - SURVEY.md section 5):
superblock v0/v1, version-1 object headers with continuation blocks, old-style groups (symbol-table message ->
v1 B-tree ``TREE`` -> ``SNOD`` leaves, names in a local ``HEAP``), dataspace v1/v2, contiguous and compact data
layout v3 (no chunking, no filters), datatypes: fixed-point, IEEE float, fixed-length string, enum (numpy ``bool``),
variable-length strings through the global heap (``GCOL``), and attributes (message 0x0C, v1-v3).
Values come back as h5py returns them: numpy arrays, numpy scalars for scalar dataspaces, ``bytes`` for strings.
Anything outside the subset raises ``H5LiteError`` instead of guessing.
"""
from __future__ import annotations

import struct

import numpy as np

_SIGNATURE = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5LiteError(RuntimeError):
    pass


def _pad8(n: int) -> int:
    return (n + 7) & ~7


class _Datatype:
    """Decoded datatype message (0x0003): numpy dtype, or a marker for variable-length strings."""

    def __init__(self, buf: bytes, off: int) -> None:
        b0 = buf[off]
        self.cls, version = b0 & 0x0F, b0 >> 4
        bits = buf[off + 1] | (buf[off + 2] << 8) | (buf[off + 3] << 16)
        self.size = struct.unpack_from("<I", buf, off + 4)[0]
        self.vlen_string = False
        self.enum_bool = False
        order = ">" if (bits & 1) else "<"
        if self.cls == 0:                                  # fixed-point
            self.dtype = np.dtype(f"{order}{'i' if bits & 0x08 else 'u'}{self.size}")
            self.length = 8 + 4
        elif self.cls == 1:                                # floating-point
            if self.size not in (2, 4, 8):
                raise H5LiteError(f"unsupported float size {self.size}")
            self.dtype = np.dtype(f"{order}f{self.size}")
            self.length = 8 + 12
        elif self.cls == 3:                                # fixed-length string
            self.dtype = np.dtype(f"S{self.size}")
            self.length = 8
        elif self.cls == 8:                                # enumeration (numpy bool is an enum over int8)
            base = _Datatype(buf, off + 8)
            n_members = bits & 0xFFFF
            p = off + 8 + base.length
            names = []
            for _ in range(n_members):
                end = buf.index(b"\x00", p)
                names.append(buf[p:end])
                p += _pad8(end - p + 1) if version < 3 else end - p + 1
            p += n_members * base.size
            self.dtype = base.dtype
            self.enum_bool = base.size == 1 and sorted(names) == [b"FALSE", b"TRUE"]
            self.length = p - off
        elif self.cls == 9:                                # variable-length
            if (bits & 0x0F) != 1:
                raise H5LiteError("variable-length sequences are not supported (only strings)")
            base = _Datatype(buf, off + 8)
            self.vlen_string = True
            self.dtype = None
            self.length = 8 + base.length
        else:
            raise H5LiteError(f"unsupported HDF5 datatype class {self.cls}")


def _dataspace(buf: bytes, off: int) -> tuple[int, ...] | None:
    """Dataspace message (0x0001) -> shape; () for scalar, None for a null dataspace."""
    version, rank, flags = buf[off], buf[off + 1], buf[off + 2]
    if version == 1:
        p = off + 8
    elif version == 2:
        if buf[off + 3] == 2:
            return None
        p = off + 4
    else:
        raise H5LiteError(f"unsupported dataspace version {version}")
    return tuple(struct.unpack_from(f"<{rank}Q", buf, p)) if rank else ()


class _Object:
    """One object header: its messages, decoded lazily into a Group or a Dataset."""

    def __init__(self, file: "File", addr: int, name: str) -> None:
        self._file, self._addr, self.name = file, addr, name
        self._messages = file._read_messages(addr)
        self.attrs = file._attributes(self._messages)

    def __bool__(self) -> bool:   # h5py objects are truthy while open (the reference relies on `if dataset`)
        return True


class Dataset(_Object):
    def __init__(self, file: "File", addr: int, name: str) -> None:
        super().__init__(file, addr, name)
        buf = file._buf
        self._dt = self._shape = self._layout = None
        for mtype, off, size in self._messages:
            if mtype == 0x0001:
                self._shape = _dataspace(buf, off)
            elif mtype == 0x0003:
                self._dt = _Datatype(buf, off)
            elif mtype == 0x0008:
                self._layout = off
            elif mtype == 0x000B:
                raise H5LiteError(f"{name}: filtered datasets are not supported")
        if self._dt is None or self._layout is None:
            raise H5LiteError(f"{name}: not a dataset")

    @property
    def shape(self) -> tuple[int, ...]:
        return self._shape or ()

    @property
    def dtype(self):
        return np.dtype(object) if self._dt.vlen_string else (np.dtype(bool) if self._dt.enum_bool else self._dt.dtype)

    def _raw(self) -> bytes:
        buf, off = self._file._buf, self._layout
        version, lclass = buf[off], buf[off + 1]
        if version != 3:
            raise H5LiteError(f"{self.name}: data layout version {version} is not supported")
        count = int(np.prod(self.shape)) if self.shape else 1
        nbytes = count * self._dt.size
        if lclass == 0:                                    # compact
            size = struct.unpack_from("<H", buf, off + 2)[0]
            return bytes(buf[off + 4:off + 4 + min(size, nbytes)])
        if lclass == 1:                                    # contiguous
            addr, size = struct.unpack_from("<QQ", buf, off + 2)
            if addr == _UNDEF:
                return b"\x00" * nbytes                    # never written: fill value 0
            return bytes(buf[addr:addr + nbytes])
        raise H5LiteError(f"{self.name}: chunked datasets are not supported")

    def __getitem__(self, key):
        if key != () and key is not Ellipsis:
            return self[()][key]
        if self._shape is None:
            raise H5LiteError(f"{self.name}: null dataspace")
        raw = self._raw()
        if self._dt.vlen_string:
            count = int(np.prod(self.shape)) if self.shape else 1
            out = [self._file._global_heap_object(*struct.unpack_from("<IQI", raw, 16 * i)) for i in range(count)]
            if self.shape == ():
                return out[0]
            return np.array(out, dtype=object).reshape(self.shape)
        arr = np.frombuffer(raw, dtype=self._dt.dtype).reshape(self.shape)
        if self._dt.enum_bool:
            arr = arr.astype(bool)
        elif self._dt.dtype.byteorder == ">":
            arr = arr.astype(self._dt.dtype.newbyteorder("<"))
        if self.shape == ():
            return arr[()]                                 # numpy scalar (bytes for fixed-length strings)
        return arr.copy()

    def __len__(self) -> int:
        return self.shape[0]

    def __repr__(self) -> str:
        return f'<h5lite dataset "{self.name}": shape {self.shape}, type {self.dtype}>'


class Group(_Object):
    def __init__(self, file: "File", addr: int, name: str) -> None:
        super().__init__(file, addr, name)
        self._links = None
        self._symtab = None
        for mtype, off, size in self._messages:
            if mtype == 0x0011:
                self._symtab = struct.unpack_from("<QQ", file._buf, off)
            elif mtype in (0x0002, 0x0006):
                raise H5LiteError(f"{name}: new-style (link message) groups are not supported")
        if self._symtab is None:
            raise H5LiteError(f"{name}: not a group")

    def _entries(self) -> dict[str, int]:
        if self._links is None:
            btree, heap = self._symtab
            self._links = dict(sorted(self._file._walk_group_btree(btree, self._file._local_heap_data(heap)).items()))
        return self._links

    def keys(self):
        return self._entries().keys()

    def __iter__(self):
        return iter(self._entries())

    def __len__(self) -> int:
        return len(self._entries())

    def __contains__(self, key: str) -> bool:
        return self.get(key) is not None

    def get(self, key: str, default=None):
        node = self
        for part in [p for p in key.split("/") if p]:
            if not isinstance(node, Group) or part not in node._entries():
                return default
            node = node._file._open(node._entries()[part], f"{node.name.rstrip('/')}/{part}")
        return node

    def __getitem__(self, key: str):
        obj = self.get(key)
        if obj is None:
            raise KeyError(f"Unable to open object (object '{key}' doesn't exist in '{self.name}')")
        return obj

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __repr__(self) -> str:
        return f'<h5lite group "{self.name}" ({len(self)} members)>'


class File(Group):
    """``with File(path) as f:`` - read-only; the whole file (scenario files are < 1 MB) is held in memory."""

    def __init__(self, path, mode: str = "r") -> None:
        if mode != "r":
            raise H5LiteError("h5lite is read-only")
        self.filename = str(path)
        with open(path, "rb") as fh:
            self._buf = fh.read()
        if self._buf[:8] != _SIGNATURE:
            raise H5LiteError(f"{path}: not an HDF5 file (or it has a user block)")
        version = self._buf[8]
        if version > 1:
            raise H5LiteError(f"{path}: superblock version {version} is not supported (written with libver='latest'?)")
        if self._buf[13] != 8 or self._buf[14] != 8:
            raise H5LiteError("only 8-byte offsets and lengths are supported")
        p = 24 if version == 0 else 28                     # v1 adds indexed-storage K + reserved
        base, _free, _eof, _driver = struct.unpack_from("<QQQQ", self._buf, p)
        if base != 0:
            raise H5LiteError("non-zero base address")
        _name_off, root_header = struct.unpack_from("<QQ", self._buf, p + 32)
        self._cache: dict[int, _Object] = {}
        super().__init__(self, root_header, "/")

    # ---- context manager / h5py.File surface --------------------------------------------------------
    def __enter__(self) -> "File":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def close(self) -> None:
        pass

    # ---- low-level pieces ---------------------------------------------------------------------------
    def _open(self, addr: int, name: str) -> _Object:
        obj = self._cache.get(addr)
        if obj is None:
            types = {m[0] for m in self._read_messages(addr)}
            obj = Group(self, addr, name) if 0x0011 in types else Dataset(self, addr, name)
            self._cache[addr] = obj
        return obj

    def _read_messages(self, addr: int) -> list[tuple[int, int, int]]:
        """Version-1 object header -> [(message type, offset of its data, size)], following continuation blocks."""
        buf = self._buf
        if buf[addr:addr + 4] == b"OHDR":
            raise H5LiteError("version-2 object headers are not supported")
        version, _, n_messages, _refs, header_size = struct.unpack_from("<BBHII", buf, addr)
        if version != 1:
            raise H5LiteError(f"object header version {version} at {addr} is not supported")
        blocks = [(addr + 16, header_size)]
        out = []
        while blocks and len(out) < n_messages:
            p, remaining = blocks.pop(0)
            end = p + remaining
            while p + 8 <= end and len(out) < n_messages:
                mtype, size, _flags = struct.unpack_from("<HHB", buf, p)
                data = p + 8
                if mtype == 0x0010:
                    blocks.append(struct.unpack_from("<QQ", buf, data))
                out.append((mtype, data, size))
                p = data + size
        return out

    def _attributes(self, messages) -> dict:
        buf, out = self._buf, {}
        for mtype, off, size in messages:
            if mtype != 0x000C:
                continue
            version = buf[off]
            name_size, dt_size, ds_size = struct.unpack_from("<HHH", buf, off + 2)
            p = off + 8 + (1 if version == 3 else 0)
            pad = _pad8 if version == 1 else (lambda n: n)
            name = buf[p:p + name_size].split(b"\x00")[0].decode("utf-8")
            p += pad(name_size)
            dt = _Datatype(buf, p)
            p += pad(dt_size)
            shape = _dataspace(buf, p)
            p += pad(ds_size)
            count = int(np.prod(shape)) if shape else 1
            if dt.vlen_string:
                vals = [self._global_heap_object(*struct.unpack_from("<IQI", buf, p + 16 * i)).decode("utf-8") for i in range(count)]
                out[name] = vals[0] if shape == () else np.array(vals, dtype=object).reshape(shape)
            else:
                arr = np.frombuffer(buf, dtype=dt.dtype, count=count, offset=p).reshape(shape or ())
                if dt.cls == 3:
                    out[name] = arr[()].decode("utf-8") if shape == () else arr.copy()
                else:
                    out[name] = arr[()] if shape == () else arr.copy()
        return out

    def _local_heap_data(self, addr: int) -> int:
        if self._buf[addr:addr + 4] != b"HEAP":
            raise H5LiteError(f"bad local heap at {addr}")
        return struct.unpack_from("<Q", self._buf, addr + 24)[0]

    def _walk_group_btree(self, addr: int, heap_data: int) -> dict[str, int]:
        buf = self._buf
        if buf[addr:addr + 4] == b"SNOD":
            n = struct.unpack_from("<H", buf, addr + 6)[0]
            links = {}
            for i in range(n):
                name_off, header = struct.unpack_from("<QQ", buf, addr + 8 + 40 * i)
                start = heap_data + name_off
                links[buf[start:buf.index(b"\x00", start)].decode("utf-8")] = header
            return links
        if buf[addr:addr + 4] != b"TREE":
            raise H5LiteError(f"bad group B-tree node at {addr}")
        node_type, _level, used = struct.unpack_from("<BBH", buf, addr + 4)
        if node_type != 0:
            raise H5LiteError("not a group B-tree")
        links = {}
        p = addr + 24                                      # key 0, then (child, key) pairs
        for i in range(used):
            child = struct.unpack_from("<Q", buf, p + 8 + 16 * i)[0]
            links.update(self._walk_group_btree(child, heap_data))
        return links

    def _global_heap_object(self, length: int, collection: int, index: int) -> bytes:
        buf = self._buf
        if length == 0:
            return b""
        if buf[collection:collection + 4] != b"GCOL":
            raise H5LiteError(f"bad global heap collection at {collection}")
        size = struct.unpack_from("<Q", buf, collection + 8)[0]
        p, end = collection + 16, collection + size
        while p + 16 <= end:
            idx, _refs, _, osize = struct.unpack_from("<HHIQ", buf, p)
            if idx == index:
                return bytes(buf[p + 16:p + 16 + osize])
            if idx == 0:
                break
            p += 16 + _pad8(osize)
        raise H5LiteError(f"global heap object {index} not found in collection at {collection}")
