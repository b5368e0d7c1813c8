"""
Minimal read-only HDF5 parser (pure Python + numpy + zlib).

h5py/libhdf5 are not available in this image, but the hot path needs two kinds of HDF5 files:
Keras model files (reference `classify.py:86-103 load_trained_model`) and fast5 reads (reference
`load_fast5s.py:25-49 get_read_id_and_signal`, `:93-98 get_root_level_keys`).  This module
implements exactly the subset of the format those files use:

  * superblock v0/v1, 8-byte offsets and lengths
  * object headers v1 (with continuation blocks) and v2 ('OHDR'/'OCHK')
  * old-style groups (symbol table message -> B-tree v1 -> SNOD -> local heap)
  * new-style groups with compact link messages, and dense link storage (fractal heap direct
    blocks are scanned for link messages; enough for listing the root keys of multi-read fast5s)
  * datasets: contiguous, compact and chunked (B-tree v1) layouts; deflate + shuffle filters
  * attributes (message v1/v2/v3): fixed-length strings, variable-length strings (global heap),
    integers and floats, scalar or simple dataspaces

It is host-side I/O only; nothing here touches the GPU.
"""

import struct
import zlib

import numpy as np

_SIGNATURE = b'\x89HDF\r\n\x1a\n'
_UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(OSError):
    """Raised for anything that is not a readable HDF5 file (mirrors h5py raising OSError)."""


class _Datatype:
    __slots__ = ('cls', 'size', 'np_dtype', 'is_vlen_string', 'base', 'encoded_len')

    def __init__(self):
        self.cls = None
        self.size = 0
        self.np_dtype = None
        self.is_vlen_string = False
        self.base = None
        self.encoded_len = 0


def _parse_datatype(buf, pos):
    dt = _Datatype()
    b0 = buf[pos]
    dt.cls = b0 & 0x0F
    bits0, bits1, bits2 = buf[pos + 1], buf[pos + 2], buf[pos + 3]
    dt.size = struct.unpack_from('<I', buf, pos + 4)[0]
    p = pos + 8
    if dt.cls == 0:  # fixed-point
        endian = '>' if (bits0 & 1) else '<'
        signed = bool(bits0 & 0x08)
        dt.np_dtype = np.dtype('{}{}{}'.format(endian, 'i' if signed else 'u', dt.size))
        p += 4
    elif dt.cls == 1:  # floating point
        endian = '>' if (bits0 & 1) else '<'
        dt.np_dtype = np.dtype('{}f{}'.format(endian, dt.size))
        p += 12
    elif dt.cls == 3:  # fixed-length string
        dt.np_dtype = np.dtype('S{}'.format(dt.size))
    elif dt.cls == 9:  # variable length
        vtype = bits0 & 0x0F
        dt.is_vlen_string = (vtype == 1)
        dt.base, p = _parse_datatype(buf, p)
    elif dt.cls == 8:  # enum (used by some fast5 attrs) - treat as base integer
        base, p2 = _parse_datatype(buf, p)
        dt.np_dtype = base.np_dtype
        # skip names/values - we never need them; encoded length is supplied by the caller
        p = p2
    elif dt.cls == 6:  # compound - not needed; caller will skip using the message size
        pass
    else:
        pass
    dt.encoded_len = p - pos
    _ = (bits1, bits2)
    return dt, p


def _parse_dataspace(buf, pos):
    version = buf[pos]
    rank = buf[pos + 1]
    flags = buf[pos + 2]
    if version == 1:
        p = pos + 8
    elif version == 2:
        p = pos + 4
    else:
        raise Hdf5Error('unsupported dataspace version {}'.format(version))
    dims = struct.unpack_from('<{}Q'.format(rank), buf, p) if rank else ()
    p += 8 * rank
    if flags & 1:
        p += 8 * rank
    return tuple(int(d) for d in dims), p


class _Message:
    __slots__ = ('type', 'data', 'flags')

    def __init__(self, mtype, data, flags):
        self.type = mtype
        self.data = data
        self.flags = flags


class Hdf5Object:
    """A group or dataset, addressed by its object-header offset."""

    def __init__(self, hfile, addr, name='/'):
        self._f = hfile
        self._addr = addr
        self.name = name
        self._messages = hfile._read_object_header(addr)
        self._links = None
        self._attrs = None

    # -- classification ---------------------------------------------------------------------
    @property
    def is_dataset(self):
        return any(m.type == 0x08 for m in self._messages)

    # -- group interface ---------------------------------------------------------------------
    def _load_links(self):
        if self._links is not None:
            return self._links
        links = {}
        f = self._f
        for m in self._messages:
            if m.type == 0x11:  # symbol table (old-style group)
                btree_addr, heap_addr = struct.unpack_from('<QQ', m.data, 0)
                heap_data_addr = f._local_heap_data_addr(heap_addr)
                f._walk_group_btree(btree_addr, heap_data_addr, links)
            elif m.type == 0x06:  # link message (compact new-style group)
                name, addr = f._parse_link(m.data, 0)[:2]
                if name is not None and addr is not None:
                    links[name] = addr
            elif m.type == 0x02:  # link info -> maybe dense storage
                d = m.data
                flags = d[1]
                p = 2
                if flags & 1:
                    p += 8
                fheap_addr, name_idx_addr = struct.unpack_from('<QQ', d, p)
                if fheap_addr != _UNDEF:
                    f._scan_fractal_heap_links(fheap_addr, links, name_idx_addr)
        self._links = links
        return links

    def keys(self):
        return list(self._load_links().keys())

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def values(self):
        return [self[k] for k in self.keys()]

    def __getitem__(self, path):
        obj = self
        for part in [p for p in path.split('/') if p]:
            links = obj._load_links()
            if part not in links:
                raise KeyError(path)
            obj = Hdf5Object(self._f, links[part], part)
        return obj

    # -- attributes --------------------------------------------------------------------------
    @property
    def attrs(self):
        if self._attrs is None:
            attrs = {}
            for m in self._messages:
                if m.type == 0x0C:
                    name, value = self._f._parse_attribute(m.data)
                    attrs[name] = value
            self._attrs = attrs
        return self._attrs

    # -- dataset interface -------------------------------------------------------------------
    def read(self):
        """Return the whole dataset as a numpy array (h5py's `dataset[:]`)."""
        dtype = shape = layout = None
        filters = []
        for m in self._messages:
            if m.type == 0x03:
                dtype, _ = _parse_datatype(m.data, 0)
            elif m.type == 0x01:
                shape, _ = _parse_dataspace(m.data, 0)
            elif m.type == 0x08:
                layout = m.data
            elif m.type == 0x0B:
                filters = self._f._parse_filter_pipeline(m.data)
        if dtype is None or shape is None or layout is None:
            raise Hdf5Error('{} is not a dataset'.format(self.name))
        if dtype.np_dtype is None:
            raise Hdf5Error('unsupported dataset datatype class {}'.format(dtype.cls))
        return self._f._read_dataset(dtype, shape, layout, filters)

    @property
    def shape(self):
        for m in self._messages:
            if m.type == 0x01:
                return _parse_dataspace(m.data, 0)[0]
        raise Hdf5Error('no dataspace')


class Hdf5File(Hdf5Object):
    """Read-only HDF5 file.  Use as a context manager or call close()."""

    def __init__(self, path):
        try:
            with open(str(path), 'rb') as fh:
                self._buf = fh.read()
        except (IOError, OSError) as e:
            raise Hdf5Error(str(e))
        buf = self._buf
        if len(buf) < 96 or buf[:8] != _SIGNATURE:
            raise Hdf5Error('{}: not an HDF5 file'.format(path))
        version = buf[8]
        if version in (0, 1):
            if buf[13] != 8 or buf[14] != 8:
                raise Hdf5Error('only 8-byte offsets/lengths are supported')
            p = 24 if version == 0 else 28
            self._base = struct.unpack_from('<Q', buf, p)[0]
            root_entry = p + 32
            root_addr = struct.unpack_from('<Q', buf, root_entry + 8)[0]
        elif version in (2, 3):
            if buf[9] != 8 or buf[10] != 8:
                raise Hdf5Error('only 8-byte offsets/lengths are supported')
            self._base = struct.unpack_from('<Q', buf, 12)[0]
            root_addr = struct.unpack_from('<Q', buf, 36)[0]
        else:
            raise Hdf5Error('unsupported superblock version {}'.format(version))
        self._gcol_cache = {}
        Hdf5Object.__init__(self, self, root_addr, '/')

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        self._buf = b''

    # -- object headers ------------------------------------------------------------------------
    def _read_object_header(self, addr):
        buf = self._buf
        if addr + 16 > len(buf):
            raise Hdf5Error('object header out of range')
        if buf[addr:addr + 4] == b'OHDR':
            return self._read_object_header_v2(addr)
        if buf[addr] != 1:
            raise Hdf5Error('unsupported object header version {}'.format(buf[addr]))
        nmsgs = struct.unpack_from('<H', buf, addr + 2)[0]
        hdr_size = struct.unpack_from('<I', buf, addr + 8)[0]
        blocks = [(addr + 16, hdr_size)]
        messages = []
        while blocks and len(messages) < nmsgs:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(messages) < nmsgs:
                mtype, msize, mflags = struct.unpack_from('<HHB', buf, p)
                data = buf[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:
                    caddr, clen = struct.unpack_from('<QQ', data, 0)
                    blocks.append((caddr, clen))
                messages.append(_Message(mtype, data, mflags))
        return messages

    def _read_object_header_v2(self, addr):
        buf = self._buf
        flags = buf[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        size_bytes = 1 << (flags & 3)
        chunk0 = int.from_bytes(buf[p:p + size_bytes], 'little')
        p += size_bytes
        track_order = bool(flags & 0x04)
        messages = []
        blocks = [(p, chunk0)]
        while blocks:
            p, size = blocks.pop(0)
            end = p + size
            while p + 4 <= end:
                mtype = buf[p]
                msize, mflags = struct.unpack_from('<HB', buf, p + 1)
                p += 4
                if track_order:
                    p += 2
                data = buf[p:p + msize]
                p += msize
                if mtype == 0x10:
                    caddr, clen = struct.unpack_from('<QQ', data, 0)
                    # continuation chunk: 'OCHK' signature, messages, 4-byte checksum
                    blocks.append((caddr + 4, clen - 8))
                elif mtype != 0:
                    messages.append(_Message(mtype, data, mflags))
        return messages

    # -- old-style groups ------------------------------------------------------------------------
    def _local_heap_data_addr(self, heap_addr):
        buf = self._buf
        if buf[heap_addr:heap_addr + 4] != b'HEAP':
            raise Hdf5Error('bad local heap')
        return struct.unpack_from('<Q', buf, heap_addr + 24)[0]

    def _cstring(self, addr):
        end = self._buf.index(b'\x00', addr)
        return self._buf[addr:end].decode('utf-8', 'replace')

    def _walk_group_btree(self, addr, heap_data_addr, links):
        buf = self._buf
        sig = buf[addr:addr + 4]
        if sig == b'TREE':
            level = buf[addr + 5]
            nentries = struct.unpack_from('<H', buf, addr + 6)[0]
            p = addr + 24
            for i in range(nentries):
                child = struct.unpack_from('<Q', buf, p + 8)[0]
                p += 16
                self._walk_group_btree(child, heap_data_addr, links)
            _ = level
        elif sig == b'SNOD':
            nsyms = struct.unpack_from('<H', buf, addr + 6)[0]
            p = addr + 8
            for i in range(nsyms):
                name_off, obj_addr = struct.unpack_from('<QQ', buf, p)
                links[self._cstring(heap_data_addr + name_off)] = obj_addr
                p += 40
        else:
            raise Hdf5Error('bad group B-tree node')

    # -- new-style links -------------------------------------------------------------------------
    def _parse_link(self, d, p):
        """Parse one link message body at d[p:]; returns (name, addr, end_pos)."""
        version = d[p]
        flags = d[p + 1]
        if version != 1:
            return None, None, p
        q = p + 2
        link_type = 0
        if flags & 0x08:
            link_type = d[q]
            q += 1
        if flags & 0x04:
            q += 8
        if flags & 0x10:
            q += 1
        nlen_size = 1 << (flags & 3)
        nlen = int.from_bytes(d[q:q + nlen_size], 'little')
        q += nlen_size
        name = bytes(d[q:q + nlen]).decode('utf-8', 'replace')
        q += nlen
        addr = None
        if link_type == 0:
            addr = struct.unpack_from('<Q', d, q)[0]
            q += 8
        elif link_type == 1:  # soft link
            slen = struct.unpack_from('<H', d, q)[0]
            q += 2 + slen
        return name, addr, q

    def _scan_fractal_heap_links(self, fheap_addr, links, name_index_addr=_UNDEF):
        """Dense link storage: collect the link messages stored as managed objects in a fractal
        heap.  Direct blocks reachable from the root (root direct block, or one level of indirect
        block) are located; if the v2 B-tree name index is a single leaf its heap IDs select the
        live objects, otherwise the direct blocks are scanned linearly (may include stale links
        left in free space - harmless for listing `read_*` keys)."""
        buf = self._buf
        if buf[fheap_addr:fheap_addr + 4] != b'FRHP':
            raise Hdf5Error('bad fractal heap')
        p = fheap_addr + 5
        heap_id_len, io_filter_len, flags = struct.unpack_from('<HHB', buf, p)
        p += 5
        max_managed_obj = struct.unpack_from('<I', buf, p)[0]
        p += 4          # max size of managed objects
        p += 8 * 2      # next huge id, huge btree addr
        p += 8 * 2      # free space amount, free space manager addr
        p += 8 * 4      # managed space, allocated managed, iterator offset, n managed objects
        p += 8 * 4      # huge size/count, tiny size/count
        table_width = struct.unpack_from('<H', buf, p)[0]
        p += 2
        start_block_size, max_direct_size = struct.unpack_from('<QQ', buf, p)
        p += 16
        max_heap_bits, start_rows = struct.unpack_from('<HH', buf, p)
        p += 4
        root_addr = struct.unpack_from('<Q', buf, p)[0]
        p += 8
        cur_rows = struct.unpack_from('<H', buf, p)[0]
        blk_off_bytes = (max_heap_bits + 7) // 8
        checksummed = bool(flags & 0x02)
        _ = (heap_id_len, start_rows)
        if root_addr == _UNDEF:
            return

        blocks = []  # (heap offset of block start, file addr, size)

        def add_direct(addr, size):
            if addr == _UNDEF or addr + size > len(buf) or buf[addr:addr + 4] != b'FHDB':
                return
            boff = int.from_bytes(buf[addr + 13:addr + 13 + blk_off_bytes], 'little')
            blocks.append((boff, addr, size))

        if cur_rows == 0:
            add_direct(root_addr, start_block_size)
        else:
            if buf[root_addr:root_addr + 4] != b'FHIB':
                raise Hdf5Error('bad fractal heap indirect block')
            q = root_addr + 5 + 8 + blk_off_bytes
            max_direct_rows = 2
            size = start_block_size
            while size < max_direct_size:
                size *= 2
                max_direct_rows += 1
            for row in range(min(cur_rows, max_direct_rows)):
                row_size = start_block_size * (1 if row < 2 else (1 << (row - 1)))
                for _col in range(table_width):
                    child = struct.unpack_from('<Q', buf, q)[0]
                    q += 8
                    if io_filter_len:
                        q += 12
                    add_direct(child, row_size)

        hdr_len = 5 + 8 + blk_off_bytes + (4 if checksummed else 0)

        # Preferred: heap IDs from a single-leaf v2 B-tree name index.
        ids = self._btree_v2_leaf_heap_ids(name_index_addr)
        if ids is not None:
            len_bytes = min((max_direct_size.bit_length() + 7) // 8,
                            (max_managed_obj.bit_length() + 7) // 8)
            for hid in ids:
                if (hid[0] >> 4) & 3 != 0:   # not a managed object
                    continue
                off = int.from_bytes(hid[1:1 + blk_off_bytes], 'little')
                length = int.from_bytes(hid[1 + blk_off_bytes:1 + blk_off_bytes + len_bytes],
                                        'little')
                for boff, addr, size in blocks:
                    if boff <= off < boff + size:
                        q = addr + (off - boff)
                        name, oaddr, _ = self._parse_link(buf[q:q + length], 0)
                        if name is not None and oaddr is not None:
                            links[name] = oaddr
                        break
            return

        for boff, addr, size in blocks:
            q = addr + hdr_len
            end = addr + size
            while q + 10 < end:
                if buf[q] != 1:
                    break
                try:
                    name, oaddr, q2 = self._parse_link(buf, q)
                except (struct.error, IndexError):
                    break
                if name is None or q2 <= q or q2 > end:
                    break
                if oaddr is not None:
                    links[name] = oaddr
                q = q2

    def _btree_v2_leaf_heap_ids(self, addr):
        """Heap IDs of a 'link name' v2 B-tree whose root is a leaf; None if not applicable."""
        buf = self._buf
        if addr == _UNDEF or buf[addr:addr + 4] != b'BTHD':
            return None
        btype = buf[addr + 5]
        rec_size, depth = struct.unpack_from('<HH', buf, addr + 10)
        root, nrec = struct.unpack_from('<QH', buf, addr + 16)
        if btype != 5 or depth != 0:
            return None
        if nrec == 0 or root == _UNDEF:
            return []
        if buf[root:root + 4] != b'BTLF':
            return None
        ids = []
        q = root + 6
        for _ in range(nrec):
            ids.append(bytes(buf[q + 4:q + rec_size]))
            q += rec_size
        return ids

    # -- attributes ------------------------------------------------------------------------------
    def _global_heap_object(self, coll_addr, index):
        buf = self._buf
        if coll_addr not in self._gcol_cache:
            if buf[coll_addr:coll_addr + 4] != b'GCOL':
                raise Hdf5Error('bad global heap collection')
            coll_size = struct.unpack_from('<Q', buf, coll_addr + 8)[0]
            objs = {}
            p = coll_addr + 16
            end = coll_addr + coll_size
            while p + 16 <= end:
                idx, _ref, _rsv, size = struct.unpack_from('<HHIQ', buf, p)
                if idx == 0:
                    break
                objs[idx] = (p + 16, size)
                p += 16 + ((size + 7) // 8) * 8
            self._gcol_cache[coll_addr] = objs
        start, size = self._gcol_cache[coll_addr][index]
        return buf[start:start + size]

    def _parse_attribute(self, d):
        version = d[0]
        name_size, dt_size, ds_size = struct.unpack_from('<HHH', d, 2)
        if version == 1:
            p = 8
            pad = lambda n: ((n + 7) // 8) * 8
        elif version == 2:
            p = 8
            pad = lambda n: n
        elif version == 3:
            p = 9
            pad = lambda n: n
        else:
            raise Hdf5Error('unsupported attribute message version {}'.format(version))
        name = bytes(d[p:p + name_size]).split(b'\x00')[0].decode('utf-8', 'replace')
        p += pad(name_size)
        dtype, _ = _parse_datatype(d, p)
        p += pad(dt_size)
        shape = _parse_dataspace(d, p)[0] if ds_size >= 4 else ()
        p += pad(ds_size)
        count = 1
        for s in shape:
            count *= s
        if dtype.cls == 9:
            if not dtype.is_vlen_string:
                return name, None
            vals = []
            for i in range(count):
                length, gaddr, gidx = struct.unpack_from('<IQI', d, p + 16 * i)
                vals.append(self._global_heap_object(gaddr, gidx)[:length] if length else b'')
            value = vals[0] if not shape else np.array(vals, dtype=object)
            return name, value
        if dtype.np_dtype is None:
            return name, None
        arr = np.frombuffer(bytes(d[p:p + count * dtype.size]), dtype=dtype.np_dtype, count=count)
        if not shape:
            value = arr[0]
            if dtype.cls == 3:
                value = bytes(value)
            return name, value
        return name, arr.reshape(shape)

    # -- datasets --------------------------------------------------------------------------------
    @staticmethod
    def _parse_filter_pipeline(d):
        version = d[0]
        nfilters = d[1]
        p = 8 if version == 1 else 2
        filters = []
        for _ in range(nfilters):
            fid = struct.unpack_from('<H', d, p)[0]
            p += 2
            if version == 1 or fid >= 256:
                name_len = struct.unpack_from('<H', d, p)[0]
                p += 2
            else:
                name_len = 0
            flags, ncv = struct.unpack_from('<HH', d, p)
            p += 4
            if version == 1:
                p += ((name_len + 7) // 8) * 8
            else:
                p += name_len
            cvals = struct.unpack_from('<{}I'.format(ncv), d, p)
            p += 4 * ncv
            if version == 1 and ncv % 2:
                p += 4
            filters.append((fid, cvals))
            _ = flags
        return filters

    def _read_dataset(self, dtype, shape, layout, filters):
        buf = self._buf
        count = 1
        for s in shape:
            count *= s
        version = layout[0]
        if version != 3:
            raise Hdf5Error('unsupported data layout version {}'.format(version))
        cls = layout[1]
        np_dtype = dtype.np_dtype
        if cls == 0:  # compact
            size = struct.unpack_from('<H', layout, 2)[0]
            raw = bytes(layout[4:4 + size])
            return np.frombuffer(raw, dtype=np_dtype, count=count).reshape(shape).copy()
        if cls == 1:  # contiguous
            addr, size = struct.unpack_from('<QQ', layout, 2)
            if addr == _UNDEF:
                return np.zeros(shape, dtype=np_dtype)
            raw = buf[addr:addr + count * dtype.size]
            return np.frombuffer(raw, dtype=np_dtype, count=count).reshape(shape).copy()
        if cls == 2:  # chunked
            ndims = layout[2]
            btree_addr = struct.unpack_from('<Q', layout, 3)[0]
            chunk_dims = struct.unpack_from('<{}I'.format(ndims), layout, 11)
            rank = ndims - 1
            out = np.zeros(shape, dtype=np_dtype)
            if btree_addr != _UNDEF and count:
                self._walk_chunk_btree(btree_addr, ndims, chunk_dims[:rank], dtype, filters, out)
            return out
        raise Hdf5Error('unsupported layout class {}'.format(cls))

    def _walk_chunk_btree(self, addr, ndims, chunk_shape, dtype, filters, out):
        buf = self._buf
        if buf[addr:addr + 4] != b'TREE':
            raise Hdf5Error('bad chunk B-tree node')
        level = buf[addr + 5]
        nentries = struct.unpack_from('<H', buf, addr + 6)[0]
        key_size = 8 + 8 * ndims
        p = addr + 24
        for _ in range(nentries):
            chunk_size, filter_mask = struct.unpack_from('<II', buf, p)
            offsets = struct.unpack_from('<{}Q'.format(ndims), buf, p + 8)
            child = struct.unpack_from('<Q', buf, p + key_size)[0]
            p += key_size + 8
            if level > 0:
                self._walk_chunk_btree(child, ndims, chunk_shape, dtype, filters, out)
                continue
            raw = buf[child:child + chunk_size]
            for i, (fid, cvals) in reversed(list(enumerate(filters))):
                if filter_mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    esize = cvals[0] if cvals else dtype.size
                    n = len(raw) // esize
                    arr = np.frombuffer(raw, dtype=np.uint8, count=n * esize)
                    raw = arr.reshape(esize, n).T.tobytes()
                elif fid == 3:
                    raw = raw[:-4]  # fletcher32 checksum trailer
                else:
                    raise Hdf5Error('unsupported HDF5 filter id {} (e.g. VBZ)'.format(fid))
            full = int(np.prod(chunk_shape))
            have = len(raw) // dtype.size
            if len(chunk_shape) == 1:
                # rank-1 fast path; also tolerates truncated edge chunks (some fast5 writers
                # store only the elements that exist instead of a full chunk)
                off = offsets[0]
                n = min(have, chunk_shape[0], out.shape[0] - off)
                if n > 0:
                    out[off:off + n] = np.frombuffer(raw, dtype=dtype.np_dtype, count=n)
                continue
            if have < full:
                raise Hdf5Error('short chunk in a multi-dimensional dataset')
            chunk = np.frombuffer(raw, dtype=dtype.np_dtype, count=full).reshape(chunk_shape)
            src = []
            dst = []
            for d, (off, csz) in enumerate(zip(offsets, chunk_shape)):
                n = min(csz, out.shape[d] - off)
                if n <= 0:
                    break
                src.append(slice(0, n))
                dst.append(slice(off, off + n))
            else:
                out[tuple(dst)] = chunk[tuple(src)]


def open_file(path):
    return Hdf5File(path)
