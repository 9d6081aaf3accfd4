/*
 * LmdbFile.hpp -- writer and reader of LMDB 0.9 data files (MDB_DATA_VERSION 1), C++17, no third-party dependency.
 *
 * The reference stores its dataset in one LMDB environment opened MDB_NOSUBDIR with named sub-databases created
 * MDB_INTEGERKEY | MDB_CREATE (DG/Util/Dataset/Dataset.cpp:8-18, 78-90) and the training side reads it with
 * py-lmdb (`Environment(path, subdir=False, max_dbs=64)`, `open_db(name, integerkey=True)`,
 * DeepestScatter_Train/LmdbDataset.py:24-47).  Neither liblmdb nor py-lmdb exists in this environment, so this file
 * implements the on-disk format directly, restated from the published LMDB 0.9.x layout (mdb.c):
 *
 *   page       4096 bytes; header = {u64 pgno; u16 pad; u16 flags; u16 lower; u16 upper} (16 bytes), for overflow pages
 *              the last 4 bytes are the page count; then u16 node offsets growing up, nodes growing down from `upper`
 *   node       {u16 lo; u16 hi; u16 flags; u16 ksize; key; data}, 2-byte aligned.  Leaf: lo|hi<<16 = data size;
 *              F_BIGDATA: data is the u64 page number of an overflow run holding the value; F_SUBDATA: data is an
 *              MDB_db record.  Branch: lo|hi<<16|flags<<32 = child page; node 0 carries an empty key
 *   MDB_db     {u32 pad; u16 flags; u16 depth; u64 branch_pages, leaf_pages, overflow_pages, entries, root} (48 bytes)
 *   meta       pages 0 and 1: header (flags P_META) + {u32 magic 0xBEEFC0DE; u32 version 1; u64 address; u64 mapsize;
 *              MDB_db free_db, main_db; u64 last_pg; u64 txnid}; the page size lives in free_db.pad, the environment
 *              flags in free_db.flags; the meta with the larger txnid is current, a commit writes page (txnid & 1)
 *   main DB    key = table name (no terminator), value = MDB_db of the table, node flag F_SUBDATA
 *   free DB    key = u64 id of the transaction that released the pages, value = {u64 n; n page numbers, descending}
 *   a value goes to overflow pages when 8 + ksize + size exceeds nodemax = ((4096 - 16) / 2 & ~1) - 2 = 2038
 *
 * FORMAT PARITY IS UNPINNED against liblmdb itself (absent here).  It is checked by an independent pure-Python page
 * parser (deepestscatter_b200/lmdb_compat.py, tests/test_lmdb.py), which also asserts the structural invariants
 * mdb.c relies on (sorted keys, >= 2 keys per branch page, node alignment, page accounting).
 *
 * Write model: append-optimised loader.  Values that need overflow pages (the 2253-byte DisneyDescriptor records) are
 * written to the file the moment they are put; small values and the key index stay in memory; commit() writes B+tree pages
 * (leaf pages filled front to back -- keys arrive sorted, as with MDB_APPEND), records the pages it replaced in the free DB
 * and flips the meta page, so a crash between commits leaves the previous snapshot intact, exactly like an aborted LMDB
 * transaction.
 * A commit is INCREMENTAL for tables that were only appended to since the last snapshot (what the collectors do,
 * Dataset.h:203-232): every completed leaf page stays where it is, only the last (partly filled) leaf, the new leaves and the
 * branch pages above them are written -- about 1/290 of the table -- so committing after every batch of 2048 records, as the
 * reference does, costs megabytes, not the whole table.  Tables that saw an overwrite or a drop are rebuilt.  Tree pages are
 * allocated from pages that neither the current nor the previous snapshot references (free-DB records of transactions older
 * than the previous one, the rule liblmdb applies when no reader is active), so repeated commits do not grow the file.
 */
#pragma once

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace dslmdb {

constexpr uint32_t MAGIC = 0xBEEFC0DEu;
constexpr uint32_t DATA_VERSION = 1;
constexpr size_t PAGEHDRSZ = 16, NODESIZE = 8;
constexpr uint16_t P_BRANCH = 0x01, P_LEAF = 0x02, P_OVERFLOW = 0x04, P_META = 0x08;
constexpr uint16_t F_BIGDATA = 0x01, F_SUBDATA = 0x02;
constexpr uint16_t MDB_INTEGERKEY = 0x08;
constexpr uint16_t MDB_NOSUBDIR = 0x4000;
constexpr uint64_t P_INVALID = ~0ull;

struct Db {
    uint32_t pad = 0;
    uint16_t flags = 0;
    uint16_t depth = 0;
    uint64_t branch_pages = 0, leaf_pages = 0, overflow_pages = 0, entries = 0, root = P_INVALID;
};
static_assert(sizeof(Db) == 48, "MDB_db is 48 bytes");

struct Meta {
    uint32_t magic = MAGIC, version = DATA_VERSION;
    uint64_t address = 0, mapsize = 0;
    Db dbs[2];
    uint64_t last_pg = 1, txnid = 0;
};
static_assert(sizeof(Meta) == 136, "MDB_meta is 136 bytes");

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

/* a value: inline bytes, or a run of overflow pages already in the file */
struct Value {
    std::vector<uint8_t> bytes;
    uint64_t ovfPage = 0;
    uint32_t size = 0;
    bool big = false;
};

class LmdbFile {
public:
    using Table = std::map<uint32_t, Value>; /* MDB_INTEGERKEY: unsigned int order (mdb_cmp_int) */

    LmdbFile(const std::string& path, bool create, bool readonly = false) : path_(path), readonly_(readonly)
    {
        fd_ = ::open(path.c_str(), readonly ? O_RDONLY : (create ? O_RDWR | O_CREAT : O_RDWR), 0664);
        if (fd_ < 0) throw Error("cannot open " + path + ": " + strerror(errno));
        struct stat st;
        if (fstat(fd_, &st) != 0) fail("fstat");
        if (st.st_size == 0) {
            if (readonly) throw Error(path + " is empty");
            initEmpty();
        } else {
            load((uint64_t)st.st_size);
        }
    }
    ~LmdbFile()
    {
        try {
            if (!readonly_) commit();
        } catch (...) {
        }
        if (fd_ >= 0) ::close(fd_);
    }
    LmdbFile(const LmdbFile&) = delete;
    LmdbFile& operator=(const LmdbFile&) = delete;

    size_t pageSize() const { return psize_; }
    uint64_t txnid() const { return meta_.txnid; }
    uint64_t lastPage() const { return nextPg_ - 1; }
    const std::map<std::string, Table>& tables() const { return tables_; }
    bool hasTable(const std::string& name) const { return tables_.count(name) != 0; }
    bool dirty() const { return dirty_; } /* something was put, created or dropped since the last commit */
    size_t count(const std::string& name) const
    {
        const auto it = tables_.find(name);
        return it == tables_.end() ? 0 : it->second.size();
    }

    /* mdb_dbi_open(name, MDB_INTEGERKEY | MDB_CREATE) */
    void createTable(const std::string& name)
    {
        requireWritable();
        if (name.empty() || name.size() > 511) throw Error("bad table name");
        if (!tables_.count(name)) {
            tables_[name];
            dirty_ = true;
        }
    }

    /* mdb_put(key = 4-byte native unsigned int) */
    void put(const std::string& table, uint32_t key, const uint8_t* data, size_t n)
    {
        requireWritable();
        if (n > 0xfffffff0u) throw Error("value too large");
        createTable(table);
        TableTree& tr = trees_[table];
        if (tr.valid && tr.covered && key <= tr.lastKey) tr.valid = false; /* not an append: this table is rebuilt by the next commit */
        Value& v = tables_[table][key];
        releaseValue(v);
        v = Value{};
        v.size = (uint32_t)n;
        if (NODESIZE + 4 + n > nodemax()) {
            v.big = true;
            v.ovfPage = writeOverflow(data, n);
        } else {
            v.bytes.assign(data, data + n);
        }
        dirty_ = true;
    }

    bool get(const std::string& table, uint32_t key, std::vector<uint8_t>& out)
    {
        const auto t = tables_.find(table);
        if (t == tables_.end()) return false;
        const auto it = t->second.find(key);
        if (it == t->second.end()) return false;
        readValue(it->second, out);
        return true;
    }

    /* mdb_drop(dbi, 0): empty the table, keep it */
    void drop(const std::string& table)
    {
        requireWritable();
        const auto t = tables_.find(table);
        if (t == tables_.end()) return;
        for (auto& kv : t->second) releaseValue(kv.second);
        t->second.clear();
        trees_[table].valid = false;
        dirty_ = true;
    }

    /* mdb_txn_commit: tree pages (incremental for append-only tables), free-DB record for the pages replaced, meta flip */
    void commit()
    {
        requireWritable();
        if (!dirty_) return;
        flushPending();
        const uint64_t txn = meta_.txnid + 1;
        /* pages no live snapshot references: free-DB records of transactions older than the previous snapshot's */
        uint64_t poolKey = 0;
        bool havePool = false;
        for (auto it = freeRecords_.begin(); it != freeRecords_.end() && it->first + 1 <= meta_.txnid;) {
            if (!havePool) {
                poolKey = it->first;
                havePool = true;
            }
            reusable_.insert(reusable_.end(), it->second.begin(), it->second.end());
            it = freeRecords_.erase(it);
        }
        std::sort(reusable_.begin(), reusable_.end(), std::greater<uint64_t>()); /* pop_back hands out the lowest page first */
        reuseTreePages_ = true;

        /* pages released by this transaction */
        std::vector<uint64_t> freed = pendingFree_;
        freed.insert(freed.end(), metaTreePages_.begin(), metaTreePages_.end());
        std::vector<uint64_t> newMetaPages;

        std::vector<Item> mainItems;
        for (const auto& t : tables_) {
            TableTree& tr = trees_[t.first];
            const Table& tab = t.second;
            if (!tr.valid) { /* overwritten, dropped or never written: rebuild from the first key */
                for (const Child& c : tr.leaves) freed.push_back(c.page);
                freed.insert(freed.end(), tr.branchPages.begin(), tr.branchPages.end());
                tr = TableTree();
                tr.valid = true;
            }
            const bool grew = tr.covered != tab.size();
            if (grew) {
                freed.insert(freed.end(), tr.branchPages.begin(), tr.branchPages.end());
                tr.branchPages.clear();
                /* reopen the last leaf: its entries are written again together with the new ones */
                Table::const_iterator from = tab.begin();
                if (!tr.leaves.empty()) {
                    uint32_t firstKey;
                    memcpy(&firstKey, tr.leaves.back().key.data(), 4);
                    from = tab.lower_bound(firstKey);
                    freed.push_back(tr.leaves.back().page);
                    tr.covered -= tr.leafCount.back();
                    tr.leaves.pop_back();
                    tr.leafCount.pop_back();
                    for (Table::const_iterator it = from; it != tab.end() && it->first <= tr.lastKey; ++it)
                        if (it->second.big) tr.overflowPages -= ovPages(it->second.size);
                }
                std::vector<Item> items;
                for (Table::const_iterator kv = from; kv != tab.end(); ++kv) {
                    Item it;
                    it.key.resize(4);
                    memcpy(it.key.data(), &kv->first, 4);
                    if (kv->second.big) {
                        it.flags = F_BIGDATA;
                        it.dataSize = kv->second.size;
                        it.data.resize(8);
                        memcpy(it.data.data(), &kv->second.ovfPage, 8);
                        tr.overflowPages += ovPages(kv->second.size);
                    } else {
                        it.dataSize = kv->second.size;
                        it.data = kv->second.bytes;
                    }
                    items.push_back(std::move(it));
                }
                std::vector<uint64_t> leafPages;
                std::vector<uint32_t> counts;
                std::vector<Child> fresh = writeLevel(items, true, leafPages, &counts);
                tr.leaves.insert(tr.leaves.end(), fresh.begin(), fresh.end());
                tr.leafCount.insert(tr.leafCount.end(), counts.begin(), counts.end());
                tr.covered = tab.size();
                tr.lastKey = tab.empty() ? 0 : tab.rbegin()->first;
                tr.db = buildBranches(tr.leaves, tr.branchPages);
                tr.db.entries = tab.size();
                tr.db.flags = MDB_INTEGERKEY;
                tr.db.overflow_pages = tr.overflowPages;
            } else if (tr.leaves.empty()) {
                tr.db = Db();
                tr.db.flags = MDB_INTEGERKEY;
            }
            Item m;
            m.key.assign(t.first.begin(), t.first.end());
            m.flags = F_SUBDATA;
            m.dataSize = sizeof(Db);
            m.data.resize(sizeof(Db));
            memcpy(m.data.data(), &tr.db, sizeof(Db));
            mainItems.push_back(std::move(m));
        }
        /* main DB keys compare as byte strings (mdb_cmp_memn) */
        std::sort(mainItems.begin(), mainItems.end(), [](const Item& a, const Item& b) {
            const size_t n = std::min(a.key.size(), b.key.size());
            const int c = memcmp(a.key.data(), b.key.data(), n);
            return c != 0 ? c < 0 : a.key.size() < b.key.size();
        });
        Db mainDb = buildTree(mainItems, newMetaPages);

        /* free DB: this transaction's record lists `freed`; what is left of the reusable pool goes back under its old id.
         * The free DB's own pages are fresh ones (their number depends on the records, which depend on the pool). */
        reuseTreePages_ = false;
        std::sort(freed.begin(), freed.end());
        freed.erase(std::unique(freed.begin(), freed.end()), freed.end());
        if (!freed.empty()) freeRecords_[txn] = freed;
        if (!reusable_.empty()) {
            std::vector<uint64_t>& back = freeRecords_[havePool ? poolKey : 0];
            back.insert(back.end(), reusable_.begin(), reusable_.end());
            std::sort(back.begin(), back.end());
            reusable_.clear();
        }
        std::vector<Item> freeItems;
        uint64_t freeOvf = 0;
        for (const auto& fr : freeRecords_) {
            Item it;
            it.key.resize(8);
            memcpy(it.key.data(), &fr.first, 8);
            std::vector<uint64_t> idl(fr.second.size() + 1);
            idl[0] = fr.second.size();
            for (size_t i = 0; i < fr.second.size(); i++) idl[i + 1] = fr.second[fr.second.size() - 1 - i]; /* descending */
            const size_t bytes = idl.size() * 8;
            it.dataSize = (uint32_t)bytes;
            if (NODESIZE + 8 + bytes > nodemax()) {
                const uint64_t pg = writeOverflow((const uint8_t*)idl.data(), bytes);
                flushPending();
                for (uint64_t p = 0; p < ovPages(bytes); p++) newMetaPages.push_back(pg + p); /* rewritten by the next commit */
                freeOvf += ovPages(bytes);
                it.flags = F_BIGDATA;
                it.data.resize(8);
                memcpy(it.data.data(), &pg, 8);
            } else {
                it.data.assign((const uint8_t*)idl.data(), (const uint8_t*)idl.data() + bytes);
            }
            freeItems.push_back(std::move(it));
        }
        Db freeDb = buildTree(freeItems, newMetaPages);
        freeDb.overflow_pages = freeOvf;
        freeDb.pad = (uint32_t)psize_;                             /* mm_psize */
        freeDb.flags = (uint16_t)(MDB_INTEGERKEY | MDB_NOSUBDIR); /* mm_flags: env flags & 0xffff | MDB_INTEGERKEY */

        flushPending();
        if (fsync(fd_) != 0) fail("fsync");
        Meta m = meta_;
        m.dbs[0] = freeDb;
        m.dbs[1] = mainDb;
        m.last_pg = nextPg_ - 1;
        m.txnid = txn;
        uint64_t need = nextPg_ * psize_;
        while (m.mapsize < need) m.mapsize *= 2; /* Dataset::increaseSize doubles the map (Dataset.cpp:55-67) */
        writeMeta(m, txn & 1);
        if (fsync(fd_) != 0) fail("fsync");
        meta_ = m;
        metaTreePages_.swap(newMetaPages);
        pendingFree_.clear();
        dirty_ = false;
    }

private:
    struct Item {
        std::vector<uint8_t> key, data; /* data: the bytes stored in the node (value, page number or MDB_db) */
        uint32_t dataSize = 0;          /* lo|hi of the node: size of the value */
        uint16_t flags = 0;
    };

    size_t nodemax() const { return (((psize_ - PAGEHDRSZ) / 2) & ~(size_t)1) - 2; }
    uint64_t ovPages(size_t n) const { return (PAGEHDRSZ - 1 + n) / psize_ + 1; }
    [[noreturn]] void fail(const char* what) const { throw Error(std::string(what) + " failed on " + path_ + ": " + strerror(errno)); }
    void requireWritable() const
    {
        if (readonly_) throw Error(path_ + " is opened read-only");
    }

    void pwriteAll(const void* buf, size_t n, uint64_t off)
    {
        const uint8_t* p = (const uint8_t*)buf;
        while (n) {
            const ssize_t w = ::pwrite(fd_, p, n, (off_t)off);
            if (w <= 0) fail("pwrite");
            p += w;
            n -= (size_t)w;
            off += (uint64_t)w;
        }
    }
    void preadAll(void* buf, size_t n, uint64_t off) const
    {
        uint8_t* p = (uint8_t*)buf;
        while (n) {
            const ssize_t r = ::pread(fd_, p, n, (off_t)off);
            if (r <= 0) throw Error("short read in " + path_);
            p += r;
            n -= (size_t)r;
            off += (uint64_t)r;
        }
    }

    /* write-combining buffer for freshly allocated pages (always a contiguous run ending at nextPg_) */
    void flushPending()
    {
        for (auto& r : reuseWrites_) pwriteAll(r.second->data(), psize_, r.first * psize_);
        reuseWrites_.clear();
        if (pendingBuf_.empty()) return;
        pwriteAll(pendingBuf_.data(), pendingBuf_.size(), pendingStart_ * psize_);
        pendingBuf_.clear();
    }
    uint8_t* allocPages(uint64_t n, uint64_t& pgno)
    {
        if (n == 1 && reuseTreePages_ && !reusable_.empty()) { /* a tree page goes to a page no live snapshot references */
            pgno = reusable_.back();
            reusable_.pop_back();
            reuseWrites_.emplace_back(pgno, std::make_unique<std::vector<uint8_t>>(psize_, (uint8_t)0));
            return reuseWrites_.back().second->data();
        }
        if (pendingBuf_.empty()) pendingStart_ = nextPg_;
        pgno = nextPg_;
        nextPg_ += n;
        const size_t at = pendingBuf_.size();
        pendingBuf_.resize(at + n * psize_, 0);
        return pendingBuf_.data() + at;
    }
    void maybeFlush()
    {
        if (pendingBuf_.size() >= (16u << 20) || reuseWrites_.size() >= 4096) flushPending();
    }

    static void putHeader(uint8_t* page, uint64_t pgno, uint16_t flags, uint16_t lower, uint16_t upper)
    {
        memcpy(page, &pgno, 8);
        const uint16_t pad = 0;
        memcpy(page + 8, &pad, 2);
        memcpy(page + 10, &flags, 2);
        memcpy(page + 12, &lower, 2);
        memcpy(page + 14, &upper, 2);
    }

    uint64_t writeOverflow(const uint8_t* data, size_t n)
    {
        const uint64_t pages = ovPages(n);
        uint64_t pgno;
        uint8_t* p = allocPages(pages, pgno);
        memcpy(p, &pgno, 8);
        const uint16_t pad = 0, flags = P_OVERFLOW;
        const uint32_t count = (uint32_t)pages;
        memcpy(p + 8, &pad, 2);
        memcpy(p + 10, &flags, 2);
        memcpy(p + 12, &count, 4);
        memcpy(p + PAGEHDRSZ, data, n);
        maybeFlush();
        return pgno;
    }

    void readValue(const Value& v, std::vector<uint8_t>& out)
    {
        if (!v.big) {
            out = v.bytes;
            return;
        }
        out.resize(v.size);
        if (!pendingBuf_.empty() && v.ovfPage >= pendingStart_) {
            memcpy(out.data(), pendingBuf_.data() + (v.ovfPage - pendingStart_) * psize_ + PAGEHDRSZ, v.size);
            return;
        }
        preadAll(out.data(), v.size, v.ovfPage * psize_ + PAGEHDRSZ);
    }

    void releaseValue(const Value& v)
    {
        if (!v.big) return;
        for (uint64_t p = 0; p < ovPages(v.size); p++) pendingFree_.push_back(v.ovfPage + p);
    }

    /* one level of pages over `items`, filled front to back; returns the (first key, page) of every page written */
    struct Child {
        std::vector<uint8_t> key;
        uint64_t page;
    };
    std::vector<Child> writeLevel(const std::vector<Item>& items, bool leaf, std::vector<uint64_t>& pagesOut, std::vector<uint32_t>* countsOut = nullptr)
    {
        /* partition.  Pages are filled the way liblmdb leaves them after inserts in ascending key order (the collectors append): when the
         * next node does not fit, mdb_page_split (newindx >= nkeys: "bias the split so the new page is emptier than the old page") moves the
         * LAST node of the full page to the new page together with the new one, so every completed page holds one node less than fit.
         * The page statistics then equal liblmdb's (DatasetVisualisation.ipynb keeps them for the authors' dataset: 84 ScatterSample records
         * per leaf where 85 fit, 290 of 291 children per branch page). */
        const size_t room = psize_ - PAGEHDRSZ;
        auto nodeNeed = [&](size_t i, bool first) {
            const size_t ksize = (!leaf && first) ? 0 : items[i].key.size(); /* a branch page's first node has an empty key */
            const size_t need = NODESIZE + ksize + (leaf ? items[i].data.size() : 0);
            return ((need + 1) & ~(size_t)1) + 2;
        };
        std::vector<size_t> starts;
        const size_t minKeep = leaf ? 1 : 2; /* mdb_page_search_root asserts NUMKEYS > 1 on branch pages */
        for (size_t i = 0; i < items.size();) {
            starts.push_back(i);
            size_t used = nodeNeed(i, true), j = i + 1;
            while (j < items.size() && used + nodeNeed(j, false) <= room) used += nodeNeed(j++, false);
            if (j < items.size() && j - i > minKeep) j -= 1;
            i = j;
        }
        /* mdb_page_search_root asserts NUMKEYS > 1 on branch pages: never leave a single node on the last page */
        if (!leaf && starts.size() >= 2 && items.size() - starts.back() < 2) starts.back() -= 1;
        std::vector<Child> out;
        for (size_t s = 0; s < starts.size(); s++) {
            const size_t b = starts[s], e = s + 1 < starts.size() ? starts[s + 1] : items.size();
            uint64_t pgno;
            uint8_t* page = allocPages(1, pgno);
            uint16_t lower = (uint16_t)PAGEHDRSZ, upper = (uint16_t)psize_;
            for (size_t i = b; i < e; i++) {
                const Item& it = items[i];
                const bool emptyKey = !leaf && i == b;
                const size_t ksize = emptyKey ? 0 : it.key.size();
                size_t nsz = NODESIZE + ksize + (leaf ? it.data.size() : 0);
                nsz = (nsz + 1) & ~(size_t)1;
                upper = (uint16_t)(upper - nsz);
                uint8_t* node = page + upper;
                uint16_t lo, hi, fl;
                if (leaf) {
                    lo = (uint16_t)(it.dataSize & 0xffffu);
                    hi = (uint16_t)(it.dataSize >> 16);
                    fl = it.flags;
                } else {
                    uint64_t child;
                    memcpy(&child, it.data.data(), 8);
                    lo = (uint16_t)(child & 0xffffu);
                    hi = (uint16_t)((child >> 16) & 0xffffu);
                    fl = (uint16_t)((child >> 32) & 0xffffu);
                }
                const uint16_t ks = (uint16_t)ksize;
                memcpy(node, &lo, 2);
                memcpy(node + 2, &hi, 2);
                memcpy(node + 4, &fl, 2);
                memcpy(node + 6, &ks, 2);
                if (ksize) memcpy(node + NODESIZE, it.key.data(), ksize);
                if (leaf && !it.data.empty()) memcpy(node + NODESIZE + ksize, it.data.data(), it.data.size());
                memcpy(page + lower, &upper, 2);
                lower = (uint16_t)(lower + 2);
            }
            if (lower > upper) throw Error("internal: page overflow");
            putHeader(page, pgno, leaf ? P_LEAF : P_BRANCH, lower, upper);
            pagesOut.push_back(pgno);
            if (countsOut) countsOut->push_back((uint32_t)(e - b));
            out.push_back(Child{items[b].key, pgno});
            maybeFlush();
        }
        return out;
    }

    Db buildTree(const std::vector<Item>& items, std::vector<uint64_t>& pagesOut)
    {
        Db db;
        db.entries = items.size();
        if (items.empty()) return db;
        std::vector<Child> level = writeLevel(items, true, pagesOut);
        Db up = buildBranches(level, pagesOut);
        up.entries = items.size();
        return up;
    }

    /* the branch levels over a list of leaf pages (first key, page), fresh pages recorded in pagesOut */
    Db buildBranches(std::vector<Child> level, std::vector<uint64_t>& pagesOut)
    {
        Db db;
        if (level.empty()) return db;
        db.leaf_pages = level.size();
        db.depth = 1;
        while (level.size() > 1) {
            std::vector<Item> up(level.size());
            for (size_t i = 0; i < level.size(); i++) {
                up[i].key = level[i].key;
                up[i].data.resize(8);
                memcpy(up[i].data.data(), &level[i].page, 8);
            }
            level = writeLevel(up, false, pagesOut);
            db.branch_pages += level.size();
            db.depth++;
        }
        db.root = level[0].page;
        flushPending();
        return db;
    }

    void writeMeta(const Meta& m, int which)
    {
        std::vector<uint8_t> page(psize_, 0);
        putHeader(page.data(), (uint64_t)which, P_META, 0, 0);
        memcpy(page.data() + PAGEHDRSZ, &m, sizeof(Meta));
        pwriteAll(page.data(), psize_, (uint64_t)which * psize_);
    }

    /* mdb_env_init_meta: both metas with txnid 0, empty trees, last page 1 */
    void initEmpty()
    {
        psize_ = 4096;
        Meta m;
        m.mapsize = 1048576; /* DEFAULT_MAPSIZE: the reference never calls mdb_env_set_mapsize before the first MapFull */
        m.dbs[0].pad = (uint32_t)psize_;
        m.dbs[0].flags = (uint16_t)(MDB_INTEGERKEY | MDB_NOSUBDIR);
        m.last_pg = 1;
        m.txnid = 0;
        writeMeta(m, 0);
        writeMeta(m, 1);
        meta_ = m;
        nextPg_ = 2;
    }

    /* ---- reading ---- */
    void readPage(uint64_t pgno, std::vector<uint8_t>& buf) const
    {
        buf.resize(psize_);
        if (pgno >= filePages_) throw Error("page number beyond the end of " + path_);
        preadAll(buf.data(), psize_, pgno * psize_);
    }

    /* depth-first walk of a tree; cb(key, flags, dataSize, node data) for every leaf node */
    void walk(uint64_t root, std::vector<uint64_t>& pages,
              const std::function<void(const uint8_t*, size_t, uint16_t, uint32_t, const uint8_t*)>& cb, int depth = 0,
              const std::function<void(uint64_t, bool, size_t)>* pageCb = nullptr) const
    {
        if (root == P_INVALID) return;
        if (depth > 32) throw Error("tree too deep (corrupt file?)");
        std::vector<uint8_t> page;
        readPage(root, page);
        uint16_t flags, lower;
        memcpy(&flags, page.data() + 10, 2);
        memcpy(&lower, page.data() + 12, 2);
        pages.push_back(root);
        const size_t n = (lower - PAGEHDRSZ) / 2;
        if (pageCb) (*pageCb)(root, (flags & P_LEAF) != 0, n);
        for (size_t i = 0; i < n; i++) {
            uint16_t off;
            memcpy(&off, page.data() + PAGEHDRSZ + 2 * i, 2);
            if ((size_t)off + NODESIZE > psize_) throw Error("node offset out of page");
            const uint8_t* node = page.data() + off;
            uint16_t lo, hi, fl, ks;
            memcpy(&lo, node, 2);
            memcpy(&hi, node + 2, 2);
            memcpy(&fl, node + 4, 2);
            memcpy(&ks, node + 6, 2);
            if (flags & P_BRANCH) {
                const uint64_t child = (uint64_t)lo | ((uint64_t)hi << 16) | ((uint64_t)fl << 32);
                walk(child, pages, cb, depth + 1, pageCb);
            } else if (flags & P_LEAF) {
                const uint32_t dsize = (uint32_t)lo | ((uint32_t)hi << 16);
                cb(node + NODESIZE, ks, fl, dsize, node + NODESIZE + ks);
            } else {
                throw Error("unexpected page type in a tree");
            }
        }
    }

    void load(uint64_t fileSize)
    {
        /* read both metas with a provisional page size, pick the newer one (mdb_env_pick_meta) */
        uint8_t head[PAGEHDRSZ + sizeof(Meta)];
        psize_ = 4096;
        Meta m[2];
        bool ok[2] = {false, false};
        for (int i = 0; i < 2; i++) {
            if (fileSize < (uint64_t)i * psize_ + sizeof(head)) break;
            preadAll(head, sizeof(head), (uint64_t)i * psize_);
            memcpy(&m[i], head + PAGEHDRSZ, sizeof(Meta));
            uint16_t pf;
            memcpy(&pf, head + 10, 2);
            ok[i] = (pf & P_META) && m[i].magic == MAGIC && m[i].version == DATA_VERSION;
            if (i == 0 && ok[0]) psize_ = m[0].dbs[0].pad;
        }
        if (!ok[0] && !ok[1]) throw Error(path_ + " is not an LMDB data file (MDB_INVALID)");
        const int cur = (ok[1] && (!ok[0] || m[1].txnid > m[0].txnid)) ? 1 : 0;
        meta_ = m[cur];
        psize_ = meta_.dbs[0].pad;
        if (psize_ < 512 || psize_ > 65536 || (psize_ & (psize_ - 1))) throw Error("unsupported page size");
        filePages_ = fileSize / psize_;
        nextPg_ = meta_.last_pg + 1;
        if (nextPg_ > filePages_) throw Error(path_ + ": last page beyond the end of the file");

        std::vector<std::pair<std::string, Db>> subs;
        walk(meta_.dbs[1].root, metaTreePages_, [&](const uint8_t* k, size_t ks, uint16_t fl, uint32_t ds, const uint8_t* d) {
            if (!(fl & F_SUBDATA) || ds != sizeof(Db)) throw Error("main DB holds a plain record: not a DeepestScatter dataset");
            Db db;
            memcpy(&db, d, sizeof(Db));
            subs.emplace_back(std::string((const char*)k, ks), db);
        });
        for (const auto& s : subs) {
            if (!(s.second.flags & MDB_INTEGERKEY)) throw Error("table " + s.first + " is not MDB_INTEGERKEY");
            Table& t = tables_[s.first];
            /* remember where the leaves are: an append-only continuation keeps them (commit()) */
            TableTree& tr = trees_[s.first];
            tr.db = s.second;
            tr.overflowPages = s.second.overflow_pages;
            std::vector<uint64_t> scratch;
            const std::function<void(uint64_t, bool, size_t)> onPage = [&](uint64_t pgno, bool leaf, size_t nkeys) {
                if (leaf) {
                    tr.leaves.push_back(Child{{}, pgno});
                    tr.leafCount.push_back((uint32_t)nkeys);
                } else {
                    tr.branchPages.push_back(pgno);
                }
            };
            walk(s.second.root, scratch, [&](const uint8_t* k, size_t ks, uint16_t fl, uint32_t ds, const uint8_t* d) {
                if (ks != 4) throw Error("table " + s.first + ": key is not a 4-byte integer");
                uint32_t key;
                memcpy(&key, k, 4);
                if (!tr.leaves.empty() && tr.leaves.back().key.empty()) tr.leaves.back().key.assign(k, k + 4);
                Value v;
                v.size = ds;
                if (fl & F_BIGDATA) {
                    v.big = true;
                    memcpy(&v.ovfPage, d, 8);
                } else {
                    v.bytes.assign(d, d + ds);
                }
                t[key] = std::move(v);
            }, 0, &onPage);
            tr.covered = t.size();
            tr.lastKey = t.empty() ? 0 : t.rbegin()->first;
            tr.valid = true;
        }
        walk(meta_.dbs[0].root, metaTreePages_, [&](const uint8_t* k, size_t ks, uint16_t fl, uint32_t ds, const uint8_t* d) {
            if (ks != 8) throw Error("free DB key is not a transaction id");
            uint64_t txn;
            memcpy(&txn, k, 8);
            std::vector<uint64_t> idl(ds / 8);
            if (fl & F_BIGDATA) {
                uint64_t pg;
                memcpy(&pg, d, 8);
                preadAll(idl.data(), ds, pg * psize_ + PAGEHDRSZ);
                for (uint64_t p = 0; p < ovPages(ds); p++) metaTreePages_.push_back(pg + p);
            } else {
                memcpy(idl.data(), d, ds);
            }
            if (idl.empty() || idl[0] != idl.size() - 1) throw Error("malformed free-list record");
            std::vector<uint64_t> asc(idl.rbegin(), idl.rend() - 1);
            freeRecords_[txn] = std::move(asc);
        });
    }

    std::string path_;
    bool readonly_ = false;
    int fd_ = -1;
    size_t psize_ = 4096;
    Meta meta_;
    uint64_t nextPg_ = 2, filePages_ = 0;
    std::map<std::string, Table> tables_;
    std::map<uint64_t, std::vector<uint64_t>> freeRecords_; /* txnid -> pages, ascending */
    /* where a table's tree lives in the current snapshot */
    struct TableTree {
        std::vector<Child> leaves;       /* leaf pages in key order: first key, page */
        std::vector<uint32_t> leafCount; /* entries per leaf */
        std::vector<uint64_t> branchPages;
        uint64_t overflowPages = 0;      /* overflow pages referenced by the covered entries */
        uint32_t lastKey = 0;            /* largest key the leaves cover */
        size_t covered = 0;              /* entries the leaves cover */
        Db db;                           /* MDB_db record of the snapshot */
        bool valid = false;              /* only appended to since the snapshot: the leaves can stay */
    };
    std::map<std::string, TableTree> trees_;
    std::vector<uint64_t> metaTreePages_; /* pages of the main DB and free DB trees (rewritten by every commit) */
    std::vector<uint64_t> reusable_;      /* single pages neither the current nor the previous snapshot references */
    std::vector<std::pair<uint64_t, std::unique_ptr<std::vector<uint8_t>>>> reuseWrites_; /* reused pages waiting for flushPending */
    bool reuseTreePages_ = false;
    std::vector<uint64_t> pendingFree_; /* overflow pages released since the last commit */
    std::vector<uint8_t> pendingBuf_;
    uint64_t pendingStart_ = 0;
    bool dirty_ = false;
};

} // namespace dslmdb
