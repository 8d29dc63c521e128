"""Minimal read-only HDF5 reader for mapped-signal files.

The reference reads its training input with h5py (taiyaki/mapped_signal_files.py:427-559);
this image has no h5py / libhdf5, so the subset of the HDF5 file format that h5py's default
writer produces for those files is decoded here in plain Python + numpy + zlib:

  * superblock version 0, version-1 object headers with continuation blocks,
  * old-style groups (symbol-table message -> version-1 B-tree -> SNOD nodes, local heap),
  * datasets: compact, contiguous and chunked layout (version-1 chunk B-tree) with the
    shuffle and deflate filters; fixed-point, floating-point, fixed strings and
    variable-length strings (global heap collections),
  * attributes (message versions 1-3) of those types.

Anything else raises `Hdf5FormatError` naming what was met.  Interface: `File(path)`,
`group[name]`, `group.keys()`, `group.attrs`, `dataset[()]` / `dataset.read()`.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5FormatError(Exception):
    pass


def _pad8(n):
    return (n + 7) & ~7


class _Datatype:
    """Decoded datatype message: `kind` in {'int', 'float', 'str', 'vlen_str', 'vlen'}."""

    def __init__(self, buf, off):
        cls_ver, b0, b1, b2, size = struct.unpack_from('<BBBBI', buf, off)
        self.cls = cls_ver & 0x0F
        self.size = size
        self.nbytes_message = 8
        if self.cls == 0:        # fixed point
            order = '>' if b0 & 1 else '<'
            signed = bool(b0 & 8)
            self.kind = 'int'
            self.dtype = np.dtype('%s%s%d' % (order, 'i' if signed else 'u', size))
            self.nbytes_message += 4
        elif self.cls == 1:      # floating point
            order = '>' if b0 & 1 else '<'
            self.kind = 'float'
            self.dtype = np.dtype('%sf%d' % (order, size))
            self.nbytes_message += 12
        elif self.cls == 3:      # fixed-length string
            self.kind = 'str'
            self.dtype = np.dtype('S%d' % size)
        elif self.cls == 9:      # variable length
            base = _Datatype(buf, off + 8)
            self.kind = 'vlen_str' if (b0 & 0x0F) == 1 else 'vlen'
            self.base = base
            self.dtype = None
            self.nbytes_message += base.nbytes_message
        else:
            raise Hdf5FormatError('datatype class %d not supported' % self.cls)


def _dataspace(buf, off):
    version, rank, flags = struct.unpack_from('<BBB', buf, off)
    if version == 1:
        p = off + 8
    elif version == 2:
        p = off + 4
        if buf[off + 3] == 2:        # null dataspace
            return None
    else:
        raise Hdf5FormatError('dataspace version %d' % version)
    return tuple(struct.unpack_from('<%dQ' % rank, buf, p)) if rank else ()


class File:
    def __init__(self, filename):
        # the file is mapped, not read: mapped-signal training files are tens of GB, and only the
        # pages of the objects that are decoded get touched
        import mmap
        with open(filename, 'rb') as fh:
            try:
                self.buf = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            except (ValueError, OSError):       # empty file / no mmap on this file system
                self.buf = fh.read()
        b = self.buf
        if b[:8] != b'\x89HDF\r\n\x1a\n':
            raise Hdf5FormatError('not an HDF5 file: %s' % filename)
        if b[8] != 0:
            raise Hdf5FormatError('superblock version %d not supported (h5py default writes 0)' % b[8])
        if b[13] != 8 or b[14] != 8:
            raise Hdf5FormatError('offset / length sizes other than 8 bytes')
        self.base = struct.unpack_from('<Q', b, 24)[0]
        root_header = struct.unpack_from('<Q', b, 56 + 8)[0]
        self._gcol = {}
        self.root = Group(self, root_header, '/')

    # convenience: behave like the root group
    def __getitem__(self, name):
        return self.root[name]

    def keys(self):
        return self.root.keys()

    @property
    def attrs(self):
        return self.root.attrs

    def close(self):
        self.buf = None

    # ---- low-level pieces
    def messages(self, addr):
        """(type, flags, offset, size) of every message of a version-1 object header."""
        b = self.buf
        if b[addr:addr + 4] == b'OHDR':
            raise Hdf5FormatError('version-2 object headers not supported')
        version, _, nmsg, _, hsize = struct.unpack_from('<BBHII', b, addr)
        if version != 1:
            raise Hdf5FormatError('object header version %d' % version)
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from('<HHB', b, p)
                body = p + 8
                if mtype == 0x10:
                    coff, clen = struct.unpack_from('<QQ', b, body)
                    blocks.append((coff + self.base, clen))
                out.append((mtype, mflags, body, msize))
                p = body + msize
        return out

    def global_heap_object(self, addr, index):
        if addr not in self._gcol:
            b = self.buf
            if b[addr:addr + 4] != b'GCOL':
                raise Hdf5FormatError('bad global heap collection at %d' % addr)
            size = struct.unpack_from('<Q', b, addr + 8)[0]
            objs = {}
            p = addr + 16
            while p + 16 <= addr + size:
                idx, _, _, osize = struct.unpack_from('<HHIQ', b, p)
                if idx == 0:
                    break
                objs[idx] = b[p + 16:p + 16 + osize]
                p += 16 + _pad8(osize)
            self._gcol[addr] = objs
        return self._gcol[addr][index]

    def decode(self, raw, dt, shape):
        """Raw element bytes -> numpy array (or list of str for variable-length strings)."""
        n = int(np.prod(shape)) if shape else 1
        if dt.kind in ('int', 'float', 'str'):
            arr = np.frombuffer(raw, dtype=dt.dtype, count=n)
            if dt.kind != 'str':
                arr = arr.astype(dt.dtype.newbyteorder('='))
            return arr.reshape(shape) if shape else arr[0]
        if dt.kind == 'vlen_str':
            out = []
            for i in range(n):
                length, gaddr, gidx = struct.unpack_from('<IQI', raw, 16 * i)
                out.append('' if gaddr == 0 and gidx == 0 else
                           self.global_heap_object(gaddr + self.base, gidx)[:length].decode('utf-8'))
            if not shape:
                return out[0]
            return np.array(out, dtype=object).reshape(shape)
        raise Hdf5FormatError('variable-length sequences not supported')


class _Node:
    def __init__(self, f, addr, name):
        self.file = f
        self.addr = addr
        self.name = name
        self._msgs = f.messages(addr)
        self._attrs = None

    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = {}
            b = self.file.buf
            for mtype, _, p, _ in self._msgs:
                if mtype != 0x0C:
                    continue
                version = b[p]
                if version == 1:
                    nsize, tsize, ssize = struct.unpack_from('<HHH', b, p + 2)
                    q = p + 8
                    name = b[q:q + nsize].split(b'\0')[0].decode()
                    q += _pad8(nsize)
                    dt = _Datatype(b, q)
                    q += _pad8(tsize)
                    shape = _dataspace(b, q)
                    q += _pad8(ssize)
                elif version in (2, 3):
                    nsize, tsize, ssize = struct.unpack_from('<HHH', b, p + 2)
                    q = p + 8 + (1 if version == 3 else 0)
                    name = b[q:q + nsize].split(b'\0')[0].decode()
                    q += nsize
                    dt = _Datatype(b, q)
                    q += tsize
                    shape = _dataspace(b, q)
                    q += ssize
                else:
                    raise Hdf5FormatError('attribute message version %d' % version)
                n = int(np.prod(shape)) if shape else 1
                raw = b[q:q + n * dt.size]
                self._attrs[name] = self.file.decode(raw, dt, shape)
        return self._attrs


class Group(_Node):
    def __init__(self, f, addr, name):
        super().__init__(f, addr, name)
        self._links = None

    def _load(self):
        if self._links is not None:
            return
        f, b = self.file, self.file.buf
        self._links = {}
        stab = [m for m in self._msgs if m[0] == 0x11]
        if not stab:
            if any(m[0] in (0x02, 0x06) for m in self._msgs):
                raise Hdf5FormatError('new-style (link message) groups not supported')
            return
        btree, heap = struct.unpack_from('<QQ', b, stab[0][2])
        heap += f.base
        if b[heap:heap + 4] != b'HEAP':
            raise Hdf5FormatError('bad local heap')
        heap_data = struct.unpack_from('<Q', b, heap + 24)[0] + f.base

        def walk(node):
            if b[node:node + 4] != b'TREE':
                raise Hdf5FormatError('bad group B-tree node')
            ntype, level, used = struct.unpack_from('<BBH', b, node + 4)
            if ntype != 0:
                raise Hdf5FormatError('group B-tree node of type %d' % ntype)
            p = node + 24
            for i in range(used):
                child = struct.unpack_from('<Q', b, p + 8)[0] + f.base
                p += 16
                if level > 0:
                    walk(child)
                    continue
                if b[child:child + 4] != b'SNOD':
                    raise Hdf5FormatError('bad symbol table node')
                nsym = struct.unpack_from('<H', b, child + 6)[0]
                for s in range(nsym):
                    e = child + 8 + 40 * s
                    name_off, header = struct.unpack_from('<QQ', b, e)
                    q = heap_data + name_off
                    name = b[q:b.find(b'\0', q)].decode()
                    self._links[name] = header + f.base
        walk(btree + f.base)

    def keys(self):
        self._load()
        return sorted(self._links)

    def __contains__(self, name):
        self._load()
        return name in self._links

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __getitem__(self, path):
        node = self
        for part in path.strip('/').split('/'):
            if not part:
                continue
            node._load()
            if part not in node._links:
                raise KeyError(part)
            addr = node._links[part]
            types = set(m[0] for m in node.file.messages(addr))
            child_name = node.name.rstrip('/') + '/' + part
            node = (Dataset if 0x08 in types else Group)(node.file, addr, child_name)
        return node


class Dataset(_Node):
    def __init__(self, f, addr, name):
        super().__init__(f, addr, name)
        b = f.buf
        self.filters = []
        for mtype, _, p, size in self._msgs:
            if mtype == 0x01:
                self.shape = _dataspace(b, p)
            elif mtype == 0x03:
                self.datatype = _Datatype(b, p)
            elif mtype == 0x08:
                self._layout = p
            elif mtype == 0x0B:
                self.filters = self._filters(p)

    def _filters(self, p):
        b = self.file.buf
        version, nfilt = b[p], b[p + 1]
        out = []
        q = p + (8 if version == 1 else 2)
        for _ in range(nfilt):
            fid = struct.unpack_from('<H', b, q)[0]
            if version == 1 or fid >= 256:
                nlen, flags, ncd = struct.unpack_from('<HHH', b, q + 2)
                q += 8
            else:
                nlen = 0
                flags, ncd = struct.unpack_from('<HH', b, q + 2)
                q += 6
            q += _pad8(nlen) if version == 1 else nlen
            cd = struct.unpack_from('<%dI' % ncd, b, q)
            q += 4 * ncd
            if version == 1 and ncd % 2:
                q += 4
            out.append((fid, cd))
        return out

    def _unfilter(self, raw, mask):
        for i in range(len(self.filters) - 1, -1, -1):
            if mask & (1 << i):
                continue
            fid, cd = self.filters[i]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                esize = cd[0] if cd else self.datatype.size
                n = len(raw) // esize
                body = np.frombuffer(raw, dtype='u1', count=n * esize).reshape(esize, n).T
                raw = body.tobytes() + raw[n * esize:]
            else:
                raise Hdf5FormatError('filter %d not supported' % fid)
        return raw

    def read(self):
        f, b, p = self.file, self.file.buf, self._layout
        dt, shape = self.datatype, self.shape
        if shape is None:
            return None
        esize = dt.size
        n = int(np.prod(shape)) if shape else 1
        version, cls = b[p], b[p + 1]
        if version != 3:
            raise Hdf5FormatError('data layout message version %d' % version)
        if cls == 0:
            size = struct.unpack_from('<H', b, p + 2)[0]
            return f.decode(b[p + 4:p + 4 + size], dt, shape)
        if cls == 1:
            addr, size = struct.unpack_from('<QQ', b, p + 2)
            if addr == UNDEF:
                return f.decode(bytes(n * esize), dt, shape)
            return f.decode(b[addr + f.base:addr + f.base + size], dt, shape)
        if cls != 2:
            raise Hdf5FormatError('data layout class %d' % cls)
        ndim = b[p + 2]
        btree = struct.unpack_from('<Q', b, p + 3)[0]
        cdims = struct.unpack_from('<%dI' % ndim, b, p + 11)
        rank = ndim - 1
        if rank != len(shape) or rank == 0:
            raise Hdf5FormatError('chunked layout of rank %d for shape %r' % (rank, shape))
        cshape = cdims[:rank]
        out = np.zeros(shape, dtype='V%d' % esize)
        if btree == UNDEF:
            return f.decode(out.tobytes(), dt, shape)

        def walk(node):
            if b[node:node + 4] != b'TREE':
                raise Hdf5FormatError('bad chunk B-tree node')
            ntype, level, used = struct.unpack_from('<BBH', b, node + 4)
            if ntype != 1:
                raise Hdf5FormatError('chunk B-tree node of type %d' % ntype)
            ksize = 8 + 8 * ndim
            q = node + 24
            for _ in range(used):
                csize, mask = struct.unpack_from('<II', b, q)
                offs = struct.unpack_from('<%dQ' % ndim, b, q + 8)[:rank]
                child = struct.unpack_from('<Q', b, q + ksize)[0] + f.base
                q += ksize + 8
                if level > 0:
                    walk(child)
                    continue
                raw = self._unfilter(b[child:child + csize], mask)
                chunk = np.frombuffer(raw, dtype='V%d' % esize,
                                      count=int(np.prod(cshape))).reshape(cshape)
                sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, shape))
                sub = tuple(slice(0, s.stop - s.start) for s in sel)
                out[sel] = chunk[sub]
        walk(btree + f.base)
        return f.decode(out.tobytes(), dt, shape)

    def __getitem__(self, key):
        val = self.read()
        if key is Ellipsis or (isinstance(key, tuple) and len(key) == 0):
            return val
        return val[key]
