"""Pure-Python reader of LMDB 0.9 data files, with the slice of the py-lmdb API that the training side of the
reference uses (DeepestScatter_Train/LmdbDataset.py:24-66): ``Environment(path, subdir=False, max_dbs=64,
readonly=True)``, ``open_db(name, integerkey=True)``, ``begin(db=..., buffers=...)``, ``Transaction.get / stat /
cursor``.  py-lmdb and liblmdb do not exist in this environment, so this module parses the on-disk format itself
(layout restated in deepestscatter_b200/host/LmdbFile.hpp, which is the writer).  It is written independently of the
C++ writer -- it walks the B+trees by binary search exactly as mdb_page_search / mdb_node_search do -- and
``check()`` asserts the structural invariants mdb.c relies on.  Read-only: ``begin(write=True)`` raises.
"""
from __future__ import annotations

import io
import mmap
import os
import struct

MAGIC = 0xBEEFC0DE
P_BRANCH, P_LEAF, P_OVERFLOW, P_META = 0x01, 0x02, 0x04, 0x08
F_BIGDATA, F_SUBDATA, F_DUPDATA = 0x01, 0x02, 0x04
MDB_INTEGERKEY = 0x08
PAGEHDRSZ, NODESIZE = 16, 8
P_INVALID = 0xFFFFFFFFFFFFFFFF

_DB = struct.Struct("<IHHQQQQQ")  # pad, flags, depth, branch, leaf, overflow, entries, root
_META = struct.Struct("<IIQQ")  # magic, version, address, mapsize


class Error(Exception):
    pass


class _Db:
    __slots__ = ("pad", "flags", "depth", "branch_pages", "leaf_pages", "overflow_pages", "entries", "root", "name")

    def __init__(self, raw, name=None):
        (self.pad, self.flags, self.depth, self.branch_pages, self.leaf_pages, self.overflow_pages, self.entries, self.root) = _DB.unpack(raw)
        self.name = name

    @property
    def integerkey(self):
        return bool(self.flags & MDB_INTEGERKEY)


class Environment:
    def __init__(self, path, map_size=10485760, subdir=True, readonly=True, max_dbs=0, mode=0o755, create=False, lock=False, **_):
        if not readonly or create:
            raise Error("lmdb_compat is a read-only implementation; datasets are written by libdeepestscatter_b200 (ds_dataset_*)")
        self._path = os.path.join(path, "data.mdb") if subdir else path
        self._f = io.open(self._path, "rb")
        size = os.fstat(self._f.fileno()).st_size
        if size < 2 * 512:
            raise Error(f"{self._path}: MDB_INVALID (file too short)")
        self._m = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        metas = []
        psize = 4096
        for i in range(2):
            off = i * psize
            if off + PAGEHDRSZ + 136 > size:
                continue
            flags = struct.unpack_from("<H", self._m, off + 10)[0]
            magic, version, _addr, mapsize = _META.unpack_from(self._m, off + PAGEHDRSZ)
            if not (flags & P_META) or magic != MAGIC or version != 1:
                continue
            free = _Db(self._m[off + PAGEHDRSZ + 24: off + PAGEHDRSZ + 72])
            main = _Db(self._m[off + PAGEHDRSZ + 72: off + PAGEHDRSZ + 120])
            last_pg, txnid = struct.unpack_from("<QQ", self._m, off + PAGEHDRSZ + 120)
            metas.append(dict(index=i, mapsize=mapsize, free=free, main=main, last_pg=last_pg, txnid=txnid, psize=free.pad))
            if i == 0:
                psize = free.pad
        if not metas:
            raise Error(f"{self._path}: MDB_INVALID (no valid meta page)")
        self._meta = max(metas, key=lambda m: (m["txnid"], -m["index"]))  # mdb_env_pick_meta: the newer one, meta 0 on a tie
        self._psize = self._meta["psize"]
        if (self._meta["last_pg"] + 1) * self._psize > size:
            raise Error(f"{self._path}: last page beyond the end of the file")
        self._max_dbs = max_dbs
        self._dbs = {}

    # ---- py-lmdb surface
    def close(self):
        if self._m is not None:
            try:
                self._m.close()
            except BufferError:  # value buffers handed out with buffers=True are still alive; the map goes with them
                pass
            self._f.close()
            self._m = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def open_db(self, key=None, txn=None, reverse_key=False, dupsort=False, create=False, integerkey=False, **_):
        if create:
            raise Error("read-only")
        if key is None:
            return self._meta["main"]
        if key in self._dbs:
            return self._dbs[key]
        if len(self._dbs) >= self._max_dbs:
            raise Error("MDB_DBS_FULL: Environment maxdbs limit reached")
        node = self._search(self._meta["main"], key)
        if node is None:
            raise Error(f"MDB_NOTFOUND: no such database {key!r}")
        flags, value = node
        if not flags & F_SUBDATA:
            raise Error("MDB_INCOMPATIBLE: not a named database")
        db = _Db(bytes(value), name=key)
        if bool(integerkey) != db.integerkey:
            raise Error("MDB_INCOMPATIBLE: integerkey flag does not match the database")
        self._dbs[key] = db
        return db

    def begin(self, db=None, parent=None, write=False, buffers=False):
        if write:
            raise Error("read-only")
        return Transaction(self, db, buffers)

    def stat(self):
        return self._stat(self._meta["main"])

    def info(self):
        m = self._meta
        return dict(map_addr=0, map_size=m["mapsize"], last_pgno=m["last_pg"], last_txnid=m["txnid"], max_readers=126, num_readers=0)

    def path(self):
        return self._path

    # ---- internals
    def _stat(self, db):
        return dict(psize=self._psize, depth=db.depth, branch_pages=db.branch_pages, leaf_pages=db.leaf_pages,
                    overflow_pages=db.overflow_pages, entries=db.entries)

    def _page(self, pgno):
        if pgno > self._meta["last_pg"]:
            raise Error(f"page {pgno} beyond last page {self._meta['last_pg']}")
        off = pgno * self._psize
        hdr_pgno, pad, flags, lower, upper = struct.unpack_from("<QHHHH", self._m, off)
        if hdr_pgno != pgno:
            raise Error(f"page {pgno}: header says {hdr_pgno}")
        return off, flags, lower, upper

    def _node(self, off, i):
        ptr = struct.unpack_from("<H", self._m, off + PAGEHDRSZ + 2 * i)[0]
        lo, hi, flags, ksize = struct.unpack_from("<HHHH", self._m, off + ptr)
        return ptr, lo, hi, flags, ksize

    def _key(self, off, ptr, ksize):
        return self._m[off + ptr + NODESIZE: off + ptr + NODESIZE + ksize]

    @staticmethod
    def _cmp(db, a, b):
        if db.integerkey:
            if len(a) != len(b):
                raise Error("MDB_BAD_VALSIZE: integer keys of different sizes")
            fmt = "<I" if len(a) == 4 else "<Q"
            x, y = struct.unpack(fmt, a)[0], struct.unpack(fmt, b)[0]
        else:
            x, y = bytes(a), bytes(b)
        return (x > y) - (x < y)

    def _leaf_value(self, off, ptr, lo, hi, flags, ksize):
        size = lo | (hi << 16)
        d = off + ptr + NODESIZE + ksize
        if flags & F_BIGDATA:
            pg = struct.unpack_from("<Q", self._m, d)[0]
            ooff, oflags, _, _ = self._page(pg)
            if not oflags & P_OVERFLOW:
                raise Error(f"page {pg} is not an overflow page")
            return memoryview(self._m)[ooff + PAGEHDRSZ: ooff + PAGEHDRSZ + size]
        return memoryview(self._m)[d: d + size]

    def _search(self, db, key):
        """mdb_page_search + mdb_node_search: returns (node flags, value view) or None."""
        pg = db.root
        if pg == P_INVALID:
            return None
        while True:
            off, flags, lower, _ = self._page(pg)
            n = (lower - PAGEHDRSZ) // 2
            if flags & P_BRANCH:
                lo_i, hi_i = 1, n - 1  # node 0 of a branch page carries the implicit lowest key
                child = 0
                while lo_i <= hi_i:
                    mid = (lo_i + hi_i) // 2
                    ptr, lo, hi, fl, ks = self._node(off, mid)
                    if self._cmp(db, key, self._key(off, ptr, ks)) >= 0:
                        child = mid
                        lo_i = mid + 1
                    else:
                        hi_i = mid - 1
                ptr, lo, hi, fl, ks = self._node(off, child)
                pg = lo | (hi << 16) | (fl << 32)
            elif flags & P_LEAF:
                lo_i, hi_i = 0, n - 1
                while lo_i <= hi_i:
                    mid = (lo_i + hi_i) // 2
                    ptr, lo, hi, fl, ks = self._node(off, mid)
                    c = self._cmp(db, key, self._key(off, ptr, ks))
                    if c == 0:
                        return fl, self._leaf_value(off, ptr, lo, hi, fl, ks)
                    if c > 0:
                        lo_i = mid + 1
                    else:
                        hi_i = mid - 1
                return None
            else:
                raise Error(f"page {pg}: unexpected flags {flags:#x} inside a tree")

    def _iterate(self, db, pages=None):
        """in-order (key, node flags, value) of a tree; `pages` collects (pgno, flags, count) of every page visited"""
        if db.root == P_INVALID:
            return
        stack = [db.root]
        while stack:
            pg = stack.pop()
            off, flags, lower, upper = self._page(pg)
            n = (lower - PAGEHDRSZ) // 2
            if pages is not None:
                pages.append((pg, flags, n, lower, upper))
            if flags & P_BRANCH:
                kids = []
                for i in range(n):
                    ptr, lo, hi, fl, ks = self._node(off, i)
                    kids.append(lo | (hi << 16) | (fl << 32))
                stack.extend(reversed(kids))
            elif flags & P_LEAF:
                for i in range(n):
                    ptr, lo, hi, fl, ks = self._node(off, i)
                    yield bytes(self._key(off, ptr, ks)), fl, self._leaf_value(off, ptr, lo, hi, fl, ks)
            else:
                raise Error(f"page {pg}: unexpected flags {flags:#x} inside a tree")


class Transaction:
    def __init__(self, env, db, buffers):
        self._env, self._db, self._buffers = env, db, buffers

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def commit(self):
        pass

    def abort(self):
        pass

    def _pick(self, db):
        db = db if db is not None else self._db
        return db if db is not None else self._env._meta["main"]

    def get(self, key, default=None, db=None):
        r = self._env._search(self._pick(db), key)
        if r is None:
            return default
        return r[1] if self._buffers else bytes(r[1])

    def stat(self, db=None):
        return self._env._stat(self._pick(db))

    def cursor(self, db=None):
        return Cursor(self, self._pick(db))

    def put(self, *a, **k):
        raise Error("read-only")


class Cursor:
    def __init__(self, txn, db):
        self._txn, self._db = txn, db
        self._it = None

    def first(self):
        self._it = self._txn._env._iterate(self._db)
        return self._db.entries > 0

    def __iter__(self):
        if self._it is None:
            self.first()
        for k, _fl, v in self._it:
            yield k, (v if self._txn._buffers else bytes(v))

    iternext = __iter__


def open(path, **kw):  # noqa: A001 - mirrors lmdb.open
    return Environment(path, **kw)


def check(path, subdir=False):
    """Structural audit of an LMDB data file.  Raises Error on the first violated invariant; returns a report."""
    env = Environment(path, subdir=subdir, readonly=True, max_dbs=1 << 20)
    m, ps = env._meta, env._psize
    nodemax = (((ps - PAGEHDRSZ) // 2) & ~1) - 2
    used = {0: "meta", 1: "meta"}
    report = dict(txnid=m["txnid"], last_pg=m["last_pg"], psize=ps, mapsize=m["mapsize"], tables={})
    if m["mapsize"] < (m["last_pg"] + 1) * ps:
        raise Error("map size smaller than the used part of the file")

    def claim(pg, what):
        if pg in used:
            raise Error(f"page {pg} used twice ({used[pg]} and {what})")
        if pg > m["last_pg"]:
            raise Error(f"page {pg} beyond last page")
        used[pg] = what

    def audit_tree(db, what, is_main=False, keysize=None):
        pages = []
        prev = None
        entries = ovf = 0
        subs = []
        for k, fl, v in env._iterate(db, pages):
            if keysize is not None and len(k) != keysize:
                raise Error(f"{what}: key of {len(k)} bytes")
            if prev is not None and Environment._cmp(db, prev, k) >= 0:
                raise Error(f"{what}: keys out of order")
            prev = k
            entries += 1
            if is_main:
                if not fl & F_SUBDATA or len(v) != 48:
                    raise Error("main DB: plain record")
                subs.append((k, _Db(bytes(v), name=k)))
        leaf = branch = 0
        depth_of = {}
        for pg, flags, n, lower, upper in pages:
            claim(pg, what)
            if lower > upper or upper > ps or lower < PAGEHDRSZ:
                raise Error(f"{what}: page {pg} lower/upper out of range")
            off = pg * ps
            if flags & P_BRANCH:
                branch += 1
                if n < 2:
                    raise Error(f"{what}: branch page {pg} has {n} keys (mdb_page_search_root asserts > 1)")
            else:
                leaf += 1
                if n < 1:
                    raise Error(f"{what}: empty leaf page {pg}")
            lowest = ps
            for i in range(n):
                ptr, lo, hi, fl, ks = env._node(off, i)
                if ptr & 1:
                    raise Error(f"{what}: page {pg} node {i} at odd offset")
                if ptr < upper or ptr + NODESIZE + ks > ps:
                    raise Error(f"{what}: page {pg} node {i} outside [upper, page end)")
                lowest = min(lowest, ptr)
                if flags & P_LEAF:
                    size = lo | (hi << 16)
                    if fl & F_BIGDATA:
                        opg = struct.unpack_from("<Q", env._m, off + ptr + NODESIZE + ks)[0]
                        ooff, oflags, _, _ = env._page(opg)
                        count = struct.unpack_from("<I", env._m, ooff + 12)[0]
                        need = (PAGEHDRSZ - 1 + size) // ps + 1
                        if not oflags & P_OVERFLOW or count != need:
                            raise Error(f"{what}: overflow run at page {opg}: flags {oflags:#x}, {count} pages, need {need}")
                        for p in range(count):
                            claim(opg + p, what + " overflow")
                        ovf += count
                    else:
                        if NODESIZE + ks + size > nodemax:
                            raise Error(f"{what}: inline node of {NODESIZE + ks + size} bytes exceeds nodemax {nodemax}")
                        if ptr + NODESIZE + ks + size > ps:
                            raise Error(f"{what}: page {pg} node {i} data beyond the page")
                elif i == 0 and ks != 0:
                    pass  # a key on branch node 0 is legal (ignored by the search)
            if n and lowest != upper:
                raise Error(f"{what}: page {pg} upper {upper} != lowest node offset {lowest}")
        if (entries, leaf, branch, ovf) != (db.entries, db.leaf_pages, db.branch_pages, db.overflow_pages):
            raise Error(f"{what}: MDB_db says entries/leaf/branch/overflow = {(db.entries, db.leaf_pages, db.branch_pages, db.overflow_pages)}, "
                        f"file holds {(entries, leaf, branch, ovf)}")
        # depth: every leaf at the same depth == db.depth
        def depth(pg, d):
            off, flags, lower, _ = env._page(pg)
            if flags & P_LEAF:
                depth_of.setdefault(d, 0)
                depth_of[d] += 1
                return
            for i in range((lower - PAGEHDRSZ) // 2):
                ptr, lo, hi, fl, ks = env._node(off, i)
                depth(lo | (hi << 16) | (fl << 32), d + 1)
        if db.root != P_INVALID:
            depth(db.root, 1)
            if list(depth_of) != [db.depth]:
                raise Error(f"{what}: leaf depths {sorted(depth_of)} but md_depth {db.depth}")
        elif db.depth != 0 or db.entries != 0:
            raise Error(f"{what}: empty root with depth {db.depth} / entries {db.entries}")
        return subs

    if m["main"].flags & MDB_INTEGERKEY:
        raise Error("main DB must not be MDB_INTEGERKEY when it holds named databases")
    subs = audit_tree(m["main"], "main", is_main=True)
    for name, db in subs:
        audit_tree(db, name.decode("ascii", "replace"), keysize=4 if db.integerkey else None)
        report["tables"][name.decode("ascii", "replace")] = env._stat(db)
    # free DB: key = txnid (8 bytes), value = {n, n page numbers descending}; freed pages must not be in use
    freed = set()
    if not m["free"].flags & MDB_INTEGERKEY:
        raise Error("free DB must be MDB_INTEGERKEY")
    for k, fl, v in env._iterate(m["free"]):
        if len(k) != 8:
            raise Error("free DB key is not a transaction id")
        txn = struct.unpack("<Q", k)[0]
        if txn > m["txnid"]:
            raise Error("free-list record from the future")
        ids = struct.unpack(f"<{len(v) // 8}Q", bytes(v))
        if not ids or ids[0] != len(ids) - 1:
            raise Error("malformed free-list record")
        if any(ids[i] <= ids[i + 1] for i in range(1, len(ids) - 1)):
            raise Error("free-list record not sorted descending")
        for pg in ids[1:]:
            if pg in freed:
                raise Error(f"page {pg} freed twice")
            freed.add(pg)
    audit_tree(m["free"], "free", keysize=8)
    both = freed & set(used)
    if both:
        raise Error(f"{len(both)} pages are both in use and on the free list, e.g. {sorted(both)[:4]}")
    total = m["last_pg"] + 1
    report.update(pages_total=total, pages_used=len(used), pages_free=len(freed), pages_leaked=total - len(used) - len(freed))
    if report["pages_leaked"] < 0:
        raise Error("page accounting negative")
    env.close()
    return report
