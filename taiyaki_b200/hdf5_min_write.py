"""Minimal writer of the HDF5 subset the batched mapped-signal format uses
(taiyaki/mapped_signal_files.py:562-679 through h5py's default settings): superblock 0,
version-1 object headers, old-style groups (B-tree + SNOD + local heap), chunked 1-D
datasets with shuffle + deflate, variable-length strings in a global heap collection,
scalar attributes.  Written from the HDF5 file-format specification, independently of
hdf5_min.py's parsing code.  It serves `mapped_signal_files.BatchHDF5Writer` and the tests
of the reader on layouts the reference's per-read fixture files do not contain.

Status: files are read back by hdf5_min.py; this image has no libhdf5 / h5py, so they have
NOT been opened with the HDF5 library itself.  Node sizes and fan-outs follow the library's
rules (symbol-table nodes of 2 x leaf-K entries, B-tree nodes padded to full size, chunk
trees of at most 64 entries per node with as many levels as needed).  Strings (read ids) go
to global heap collections of 1 MiB that are added as they fill up (at most 65535 objects
each; a string that was already stored is referenced again instead of stored twice).
Limit: at most 256 links per group."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
GCOL_BYTES = 1 << 20
CHUNK_K = 32          # chunk B-tree K (library default; superblock 0 does not store it)
GROUP_LEAF_K, GROUP_INTERNAL_K = 4, 16      # written into the superblock


def _pad8(b):
    return b + bytes(-len(b) % 8)


class Writer:
    def __init__(self):
        self.buf = bytearray(96)                 # superblock, written last
        self.collections = []                    # global heap collections: [address, objects, bytes used]
        self.string_refs = {}                    # raw string -> (collection address, object index)
        self._new_collection()

    def _new_collection(self):
        self.collections.append([self.alloc(bytes(GCOL_BYTES)), [], 16])   # 16 = collection header

    def alloc(self, data):
        self.buf += bytes(-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    # ---- messages
    @staticmethod
    def dataspace(shape):
        return struct.pack('<BBB5x', 1, len(shape), 0) + b''.join(struct.pack('<Q', s) for s in shape)

    @staticmethod
    def datatype(dtype):
        if dtype == 'vlen_str':
            base = struct.pack('<BBBBI', 0x13, 0, 0, 0, 1)             # 1-byte string
            return struct.pack('<BBBBI', 0x19, 0x01, 0x01, 0, 16) + base
        dt = np.dtype(dtype)
        if dt.kind in 'iu':
            return struct.pack('<BBBBIHH', 0x10, 0x08 if dt.kind == 'i' else 0, 0, 0, dt.itemsize,
                               0, 8 * dt.itemsize)
        if dt.kind == 'f' and dt.itemsize == 8:
            return struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, 0x3f, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
        if dt.kind == 'f' and dt.itemsize == 4:
            return struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, 0x1f, 0, 4, 0, 32, 23, 8, 0, 23, 127)
        if dt.kind == 'S':                                             # fixed-length, null-padded ASCII
            return struct.pack('<BBBBI', 0x13, 0x01, 0, 0, dt.itemsize)
        raise ValueError(dtype)

    def vlen_refs(self, values):
        """16-byte (length, collection address, index) elements."""
        out = b''
        for v in values:
            raw = v.encode('utf-8')
            ref = self.string_refs.get(raw)
            if ref is None:
                need = 16 + len(raw) + (-len(raw) % 8)               # object header + padded data
                assert need + 32 <= GCOL_BYTES, 'string longer than a heap collection'
                coll = self.collections[-1]
                # keep room for the free-space object that ends a collection
                if coll[2] + need + 16 > GCOL_BYTES or len(coll[1]) >= 65535:
                    self._new_collection()
                    coll = self.collections[-1]
                coll[1].append(raw)
                coll[2] += need
                ref = self.string_refs[raw] = (coll[0], len(coll[1]))
            out += struct.pack('<IQI', len(raw), ref[0], ref[1])
        return out

    def header(self, messages):
        body = b''
        for mtype, data in messages:
            data = _pad8(data)
            body += struct.pack('<HHB3x', mtype, len(data), 0) + data
        return self.alloc(struct.pack('<BBHII4x', 1, 0, len(messages), 1, len(body)) + body)

    def attribute(self, name, value):
        if isinstance(value, str):
            dt, data = self.datatype('vlen_str'), self.vlen_refs([value])
        else:
            arr = np.asarray(value)
            dt, data = self.datatype(arr.dtype), arr.tobytes()
        space = self.dataspace(())
        nm = name.encode() + b'\0'
        return (0x0C, struct.pack('<BxHHH', 1, len(nm), len(dt), len(space)) +
                _pad8(nm) + _pad8(dt) + _pad8(space) + data)

    # ---- objects
    def dataset(self, values, chunk, shuffle=True):
        """1-D chunked dataset, deflate (+ shuffle for numeric types)."""
        if isinstance(values, np.ndarray):
            dt_msg, esize, raw = self.datatype(values.dtype), values.dtype.itemsize, values.tobytes()
        else:
            dt_msg, esize, raw, shuffle = self.datatype('vlen_str'), 16, self.vlen_refs(values), False
        n = len(raw) // esize
        entries = []
        for start in range(0, max(n, 1), chunk):
            piece = raw[start * esize:(start + chunk) * esize]
            piece += bytes(chunk * esize - len(piece))               # edge chunks are full size
            if shuffle:
                piece = np.frombuffer(piece, dtype='u1').reshape(-1, esize).T.tobytes()
            piece = zlib.compress(piece, 4)
            entries.append((len(piece), start, self.alloc(piece)))
        btree = self._chunk_btree(entries, ((n + chunk - 1) // chunk) * chunk)
        layout = struct.pack('<BBB', 3, 2, 2) + struct.pack('<Q', btree) + struct.pack('<II', chunk, esize)
        filters = b''
        nfilt = 0
        if shuffle:
            filters += struct.pack('<HHHH', 2, 0, 1, 1) + struct.pack('<I', esize) + bytes(4)
            nfilt += 1
        filters += struct.pack('<HHHH', 1, 0, 1, 1) + struct.pack('<I', 4) + bytes(4)
        nfilt += 1
        pipeline = struct.pack('<BB6x', 1, nfilt) + filters
        return self.header([(0x01, self.dataspace((n,))), (0x03, dt_msg), (0x0B, pipeline),
                            (0x08, layout)])

    def _chunk_btree(self, entries, end_offset):
        """Version-1 B-tree (node type 1) over (stored size, first element, address) chunk
        records: at most 2K = 64 entries per node (the library's default K for chunk trees,
        implied by superblock 0), nodes padded to their full size, as many levels as needed."""
        def key(size, start):
            return struct.pack('<IIQQ', size, 0, start, 0)
        level, items = 0, [(size, start, addr) for size, start, addr in entries]
        while True:
            nodes = []
            for i in range(0, len(items), 2 * CHUNK_K):
                part = items[i:i + 2 * CHUNK_K]
                nxt = items[i + 2 * CHUNK_K][1] if i + 2 * CHUNK_K < len(items) else end_offset
                node = b'TREE' + struct.pack('<BBHQQ', 1, level, len(part), UNDEF, UNDEF)
                for size, start, addr in part:
                    node += key(size, start) + struct.pack('<Q', addr)
                node += key(0, nxt)
                node += bytes(24 + 2 * CHUNK_K * 32 + 24 - len(node))
                nodes.append((part[0][0], part[0][1], self.alloc(node)))
            if len(nodes) == 1:
                return nodes[0][2]
            level, items = level + 1, nodes

    def group(self, links, attrs=(), per_node=8):
        """Old-style group; `links` name -> object header address."""
        names = sorted(links)
        heap_data = bytearray(b'\0' * 8)
        offsets = {}
        for nm in names:
            offsets[nm] = len(heap_data)
            heap_data += _pad8(nm.encode() + b'\0')
        data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), UNDEF, data_addr))
        per_node = min(per_node, 2 * GROUP_LEAF_K)
        children = []
        for i in range(0, len(names), per_node):
            part = names[i:i + per_node]
            snod = b'SNOD' + struct.pack('<BxH', 1, len(part))
            for nm in part:
                snod += struct.pack('<QQII16x', offsets[nm], links[nm], 0, 0)
            snod += bytes(8 + 2 * GROUP_LEAF_K * 40 - len(snod))          # full-size node
            children.append((self.alloc(snod), offsets[part[-1]]))
        assert len(children) <= 2 * GROUP_INTERNAL_K, 'more than %d links in one group' % (
            4 * GROUP_LEAF_K * GROUP_INTERNAL_K)
        node = b'TREE' + struct.pack('<BBHQQ', 0, 0, len(children), UNDEF, UNDEF) + struct.pack('<Q', 0)
        for addr, last in children:
            node += struct.pack('<QQ', addr, last)
        node += bytes(24 + 2 * GROUP_INTERNAL_K * 16 + 8 - len(node))     # full-size node
        btree = self.alloc(node)
        msgs = [(0x11, struct.pack('<QQ', btree, heap))] + [self.attribute(k, v) for k, v in attrs]
        return self.header(msgs), btree, heap

    def close(self, root_header, root_btree, root_heap, filename):
        for addr, objects, used in self.collections:
            coll = b''
            for i, raw in enumerate(objects):
                coll += struct.pack('<HHIQ', i + 1, 1, 0, len(raw)) + _pad8(raw)
            head = b'GCOL' + struct.pack('<B3xQ', 1, GCOL_BYTES)
            free = GCOL_BYTES - len(head) - len(coll)
            assert len(head) + len(coll) == used and free >= 16
            coll += struct.pack('<HHIQ', 0, 0, 0, free)
            self.buf[addr:addr + len(head) + len(coll)] = head + coll
        sb = b'\x89HDF\r\n\x1a\n' + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack('<QQII', 0, root_header, 1, 0) + struct.pack('<QQ', root_btree, root_heap)
        self.buf[:96] = sb
        with open(filename, 'wb') as fh:
            fh.write(bytes(self.buf))


def write_batched_mapped_signal_file(filename, reads, batch_size=3, chunk=1000,
                                     alphabet=('ACGT', 'ACGT', '')):
    """The layout BatchHDF5Writer produces: /Batches/Batch_n/<field>[, <field>_lengths].
    `reads`: SignalMapping objects or read dictionaries; `alphabet`: (alphabet,
    collapse_alphabet, newline-joined modified-base long names)."""
    def field(r, k):
        return r[k] if isinstance(r, dict) else getattr(r, k)
    w = Writer()
    batches = {}
    for b, first in enumerate(range(0, len(reads), batch_size)):
        part = reads[first:first + batch_size]
        links = {}
        for k, dt in (('Dacs', np.int16), ('Ref_to_signal', np.int32), ('Reference', np.int16)):
            links[k] = w.dataset(np.concatenate([field(r, k) for r in part]).astype(dt), chunk)
            links[k + '_lengths'] = w.dataset(
                np.array([len(field(r, k)) for r in part], dtype=np.int32), chunk)
        for k in ('shift_frompA', 'scale_frompA', 'range', 'offset', 'digitisation'):
            links[k] = w.dataset(np.array([field(r, k) for r in part], dtype=np.float64), chunk)
        links['read_id'] = w.dataset([field(r, 'read_id') for r in part], chunk)
        batches['Batch_%d' % b] = w.group(links, per_node=5)[0]
    top = {'Batches': w.group(batches)[0]}
    if len(reads) > 0:
        top['read_ids'] = w.dataset([field(r, 'read_id') for r in reads], chunk)
    root = w.group(top, attrs=[('version', np.int64(8)), ('alphabet', alphabet[0]),
                               ('collapse_alphabet', alphabet[1]), ('mod_long_names', alphabet[2])])
    w.close(*root, filename)
