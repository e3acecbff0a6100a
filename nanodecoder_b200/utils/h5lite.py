"""The part of h5py's API that the reference's reader uses, on libnanodec's own HDF5 reader (csrc/fast5.cu: no libhdf5,
no h5py, no VBZ plugin).  The reference's `extract_fast5_raw` (utils/labelop.py:199-214) runs unmodified over it:

    import nanodecoder_b200.utils.h5lite as h5py
    fast5_data = h5py.File(path, 'r')                      # IOError (OSError) when the file is not HDF5
    raw = list(fast5_data['/Raw/Reads/'].values())[0]['Signal'].value
    fast5_data.close()

Read-only.  Groups iterate in name order like h5py's; datasets are integer or floating point, rank <= 2, any of the
layouts / filters the reader decodes (contiguous, compact, chunked with gzip / shuffle / fletcher32 / VBZ).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .labelop import _fast5_bytes


def _message(err):
    return err.value.decode("latin-1")


class _Node(object):
    def __init__(self, file, name):
        self.file = file
        self.name = name if name.startswith("/") else "/" + name

    def _path(self, key):
        if key.startswith("/"):
            return key
        return self.name.rstrip("/") + "/" + key


class Dataset(_Node):
    def __init__(self, file, name, info):
        super().__init__(file, name)
        kind, size, signed, big = int(info[0]), int(info[1]), bool(info[2]), bool(info[3])
        self.dtype = np.dtype((">" if big else "<") + ("f" if kind == 1 else "i" if signed else "u") + str(size))
        rank = int(info[4])
        if rank > 2:
            raise NotImplementedError("datasets of rank %d" % rank)
        self.shape = tuple(int(info[6 + i]) for i in range(rank))
        self._nbytes = int(info[5])

    def __len__(self):
        return self.shape[0] if self.shape else 1

    def _read(self):
        buf = (C.c_uint8 * max(1, self._nbytes))()
        info, err = (C.c_int64 * 8)(), C.create_string_buffer(512)
        rc = _lib.load().nd_h5_read_dataset(self.file._buf, self.file._n, self.name.encode("latin-1"), buf, self._nbytes,
                                            info, err, 512)
        if rc != 0:
            raise OSError("Can't read data (%s)" % _message(err))
        return np.frombuffer(bytes(buf)[: self._nbytes], self.dtype).reshape(self.shape).copy()

    @property
    def value(self):                                        # h5py < 3 (what the reference was written against)
        return self._read()

    def __getitem__(self, key):                             # ds[()], ds[...], ds[:], ds[a:b]
        data = self._read()
        return data if key is Ellipsis or key == () else data[key]

    def __array__(self, dtype=None, copy=None):
        data = self._read()
        return data if dtype is None else data.astype(dtype)


class Group(_Node):
    def keys(self):
        lib = _lib.load()
        need, count, err = C.c_int64(0), C.c_int32(0), C.create_string_buffer(512)
        path = self.name.encode("latin-1")
        rc = lib.nd_h5_list_group(self.file._buf, self.file._n, path, None, 0, C.byref(need), C.byref(count), err, 512)
        names = C.create_string_buffer(max(1, need.value))
        if rc == 0:
            rc = lib.nd_h5_list_group(self.file._buf, self.file._n, path, names, need.value, C.byref(need), C.byref(count),
                                      err, 512)
        if rc != 0:
            raise KeyError("Unable to open object (%s)" % _message(err))
        return [x.decode("latin-1") for x in names.raw[: need.value].split(b"\0")[: count.value]]

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __getitem__(self, key):
        path = self._path(key)
        info, err = (C.c_int64 * 8)(), C.create_string_buffer(512)
        rc = _lib.load().nd_h5_read_dataset(self.file._buf, self.file._n, path.encode("latin-1"), None, 0, info, err, 512)
        if rc == 0:
            return Dataset(self.file, path, info)
        why = _message(err)
        if "not a dataset" in why:
            return Group(self.file, path)
        raise KeyError("Unable to open object (%s)" % why)


class File(Group):
    def __init__(self, name, mode="r"):
        if mode not in ("r", "r+"):
            raise ValueError("h5lite opens files read-only (mode %r)" % (mode,))
        self._buf, self._n, self._keep = _fast5_bytes(name)
        self.filename = name
        Group.__init__(self, self, "/")
        try:
            self.keys()
        except KeyError as e:                               # h5py: OSError "Unable to open file (file signature not found)"
            raise OSError("Unable to open file (%s)" % e)

    def close(self):
        self._buf, self._keep = None, None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
