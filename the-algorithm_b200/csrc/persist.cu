// persist.cu -- the reference's on-disk index format, read and written natively (host code; the rows travel through the
// C ABI's own ann_read_rows / ann_append_batch).
//
//   SerializableBruteForceIndex.toDirectory   (ann/src/main/scala/com/twitter/ann/brute_force/BruteForceIndex.scala:142-161)
//   BruteForceDeserialization.fromDirectory   (ann/.../brute_force/BruteForceDeserialization.scala:42-63)
//   ThriftIteratorIO                          (ann/.../serialization/ThriftIteratorIO.scala:14-56)
//   PersistedEmbeddingInjection               (ann/.../serialization/PersistedEmbeddingInjection.scala:14-28)
//   ShardedSerialization / ComposedQueryableDeserialization (ann/.../common/ShardedSerialization.scala:17-66)
//
// One file `BruteForceFileData` per index directory: back-to-back TBinaryProtocol (big-endian, non-strict, no framing, no
// header, no count) encodings of
//     struct PersistedEmbedding { 1: required binary id; 2: required embedding.Embedding embedding }   (serialization.thrift:7-10)
// until end of file.  `id` is Injection[T, Array[Byte]] of the entity id: 8 bytes big-endian for Long (AnnInjections.scala:8),
// 4 bytes big-endian for Int (:12).  A sharded index is `shard_<i>/` sub-directories (ShardedSerialization.scala:9-11,28-38).
//
// ASSUMED LAYOUT (stated, switchable): `embedding.thrift` (com/twitter/ml/api) is NOT in the open-source tree.  The tensor
// schema it wraps IS visible through the generated code the tree ships (navi/thrift_bpr_adapter/thrift/src/tensor.rs:
// union GeneralTensor {1: RawTypedTensor, 2: StringTensor, 3: Int32Tensor, 4: Int64Tensor, 5: FloatTensor {1: list<double>
// floats, 2: optional list<i64> shape}, 6: DoubleTensor {1: list<double> doubles, 2: shape}, ...}; RawTypedTensor {1: i32
// dataType (FLOAT = 0, DOUBLE = 1), 2: binary content, 3: shape}).  The writer therefore emits
//     Embedding { 1: GeneralTensor { 5: FloatTensor { 1: list<double> } } }         layout 0 (default)
//     Embedding { 1: GeneralTensor { 6: DoubleTensor { 1: list<double> } } }        layout 1
//     Embedding { 1: GeneralTensor { 1: RawTypedTensor { 1: FLOAT, 2: <little-endian float32 bytes> } } }   layout 2
// and the READER does not depend on the choice: it walks field 2 generically and takes the first list<double> it finds at
// any depth, or a RawTypedTensor-shaped struct (i32 + binary) with dataType FLOAT / DOUBLE.
#include <cuda_runtime.h>
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200ann.h"
#include "kernels.h"

using namespace b200ann;

namespace {

constexpr const char* kDataFile = "BruteForceFileData";   // BruteForceIndex.DataFileName, BruteForceIndex.scala:27
constexpr const char* kSuccess = "_SUCCESS";              // IndexOutputFile.scala:29,58-62
constexpr const char* kShardPrefix = "shard_";            // ShardConstants.ShardPrefix, ShardedSerialization.scala:9-11

enum TType : uint8_t { T_STOP = 0, T_BOOL = 2, T_BYTE = 3, T_DOUBLE = 4, T_I16 = 6, T_I32 = 8, T_I64 = 10, T_STRING = 11,
                       T_STRUCT = 12, T_MAP = 13, T_SET = 14, T_LIST = 15 };

int perr(int code, const std::string& msg) { return report_error(code, msg.c_str()); }

// ---------------------------------------------------------------- writer
struct Writer {
    FILE* f = nullptr;
    std::vector<unsigned char> buf;
    void u8(uint8_t v) { buf.push_back(v); }
    void i16(int16_t v) { buf.push_back((uint8_t)(v >> 8)); buf.push_back((uint8_t)v); }
    void i32(int32_t v) { for (int s = 24; s >= 0; s -= 8) buf.push_back((uint8_t)((uint32_t)v >> s)); }
    void i64(int64_t v) { for (int s = 56; s >= 0; s -= 8) buf.push_back((uint8_t)((uint64_t)v >> s)); }
    void f64(double d) { uint64_t u; memcpy(&u, &d, 8); i64((int64_t)u); }
    void field(uint8_t type, int16_t id) { u8(type); i16(id); }
    bool flush() {
        const bool ok = buf.empty() || fwrite(buf.data(), 1, buf.size(), f) == buf.size();
        buf.clear();
        return ok;
    }
};

void write_record(Writer& w, int64_t id, int id_format, const float* row, int dim, int layout) {
    w.field(T_STRING, 1);                                   // 1: binary id
    if (id_format == ANN_ID_INT32_BE) { w.i32(4); w.i32((int32_t)id); }
    else { w.i32(8); w.i64(id); }
    w.field(T_STRUCT, 2);                                   // 2: embedding.Embedding
    w.field(T_STRUCT, 1);                                   //   1: tensor.GeneralTensor (union)
    if (layout == ANN_LAYOUT_RAW_FLOAT) {
        w.field(T_STRUCT, 1);                               //     1: RawTypedTensor
        w.field(T_I32, 1); w.i32(0);                        //       1: dataType = FLOAT
        w.field(T_STRING, 2); w.i32(dim * 4);               //       2: content
        const unsigned char* p = reinterpret_cast<const unsigned char*>(row);
        w.buf.insert(w.buf.end(), p, p + (size_t)dim * 4);
    } else {
        w.field(T_STRUCT, layout == ANN_LAYOUT_DOUBLE_TENSOR ? 6 : 5);   // 5: FloatTensor / 6: DoubleTensor
        w.field(T_LIST, 1); w.u8(T_DOUBLE); w.i32(dim);     //       1: list<double>
        for (int i = 0; i < dim; ++i) w.f64((double)row[i]);
    }
    w.u8(T_STOP);   // tensor struct
    w.u8(T_STOP);   // GeneralTensor
    w.u8(T_STOP);   // Embedding
    w.u8(T_STOP);   // PersistedEmbedding
}

// ---------------------------------------------------------------- reader
struct Reader {
    const unsigned char* p;
    const unsigned char* end;
    bool eof = false;        // ran off the end: the reference treats that as end of stream (ThriftIteratorIO.scala:42-49)
    bool need(size_t n) {
        if ((size_t)(end - p) < n) { eof = true; return false; }
        return true;
    }
    bool u8(uint8_t* v) { if (!need(1)) return false; *v = *p++; return true; }
    bool i16(int16_t* v) { if (!need(2)) return false; *v = (int16_t)((p[0] << 8) | p[1]); p += 2; return true; }
    bool i32(int32_t* v) {
        if (!need(4)) return false;
        *v = (int32_t)(((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]);
        p += 4;
        return true;
    }
    bool i64(int64_t* v) {
        if (!need(8)) return false;
        uint64_t u = 0;
        for (int i = 0; i < 8; ++i) u = (u << 8) | p[i];
        p += 8;
        *v = (int64_t)u;
        return true;
    }
    bool skip_bytes(size_t n) { if (!need(n)) return false; p += n; return true; }
};

struct Found {
    std::vector<float>* out;
    bool have = false;
    bool bad = false;   // malformed (negative sizes, nesting too deep)
};

bool skip_value(Reader& r, uint8_t type, int depth);

// Walk one struct.  `f` (may be null) collects the embedding: the first list<double>, or a RawTypedTensor-shaped pair.
bool walk_struct(Reader& r, Found* f, int depth) {
    if (depth > 16) { if (f) f->bad = true; return false; }
    int32_t raw_dtype = -1;
    for (;;) {
        uint8_t type;
        int16_t id;
        if (!r.u8(&type)) return false;
        if (type == T_STOP) return true;
        if (!r.i16(&id)) return false;
        if (f && !f->have && type == T_STRUCT) {
            if (!walk_struct(r, f, depth + 1)) return false;
        } else if (f && !f->have && type == T_LIST) {
            uint8_t et;
            int32_t n;
            if (!r.u8(&et) || !r.i32(&n)) return false;
            if (n < 0) { f->bad = true; return false; }
            if (et == T_DOUBLE) {
                if (!r.need((size_t)n * 8)) return false;
                f->out->resize((size_t)n);
                for (int32_t i = 0; i < n; ++i) {
                    int64_t u;
                    r.i64(&u);
                    double d;
                    memcpy(&d, &u, 8);
                    (*f->out)[(size_t)i] = (float)d;
                }
                f->have = true;
            } else {
                for (int32_t i = 0; i < n; ++i)
                    if (!skip_value(r, et, depth + 1)) return false;
            }
        } else if (f && !f->have && type == T_I32 && id == 1) {
            if (!r.i32(&raw_dtype)) return false;
        } else if (f && !f->have && type == T_STRING && id == 2 && (raw_dtype == 0 || raw_dtype == 1)) {
            int32_t n;
            if (!r.i32(&n)) return false;
            if (n < 0) { f->bad = true; return false; }
            if (!r.need((size_t)n)) return false;
            const int w = raw_dtype == 0 ? 4 : 8;
            f->out->resize((size_t)n / w);
            for (size_t i = 0; i < f->out->size(); ++i) {
                if (w == 4) memcpy(&(*f->out)[i], r.p + i * 4, 4);
                else { double d; memcpy(&d, r.p + i * 8, 8); (*f->out)[i] = (float)d; }
            }
            r.p += n;
            f->have = true;
        } else if (!skip_value(r, type, depth + 1)) {
            return false;
        }
    }
}

bool skip_value(Reader& r, uint8_t type, int depth) {
    if (depth > 32) return false;
    int32_t n;
    uint8_t a, b2;
    switch (type) {
        case T_BOOL: case T_BYTE: return r.skip_bytes(1);
        case T_I16: return r.skip_bytes(2);
        case T_I32: return r.skip_bytes(4);
        case T_I64: case T_DOUBLE: return r.skip_bytes(8);
        case T_STRING: return r.i32(&n) && n >= 0 && r.skip_bytes((size_t)n);
        case T_STRUCT: return walk_struct(r, nullptr, depth + 1);
        case T_LIST: case T_SET:
            if (!r.u8(&a) || !r.i32(&n) || n < 0) return false;
            for (int32_t i = 0; i < n; ++i)
                if (!skip_value(r, a, depth + 1)) return false;
            return true;
        case T_MAP:
            if (!r.u8(&a) || !r.u8(&b2) || !r.i32(&n) || n < 0) return false;
            for (int32_t i = 0; i < n; ++i)
                if (!skip_value(r, a, depth + 1) || !skip_value(r, b2, depth + 1)) return false;
            return true;
        default: return false;
    }
}

// one PersistedEmbedding; returns 1 = record read, 0 = clean/partial end of stream, < 0 = ann_status
int read_record(Reader& r, int id_format, int64_t* id, std::vector<float>* row) {
    if (r.p == r.end) return 0;
    bool have_id = false;
    Found f{row};
    for (;;) {
        uint8_t type;
        int16_t fid;
        if (!r.u8(&type)) return 0;
        if (type == T_STOP) break;
        if (!r.i16(&fid)) return 0;
        if (fid == 1 && type == T_STRING) {
            int32_t n;
            if (!r.i32(&n)) return 0;
            if (n < 0) return perr(ANN_ERR_INVALID_ARGUMENT, "BruteForceFileData: negative id length");
            if (!r.need((size_t)n)) return 0;
            if (n == 8 && id_format != ANN_ID_INT32_BE) {
                uint64_t u = 0;
                for (int i = 0; i < 8; ++i) u = (u << 8) | r.p[i];
                *id = (int64_t)u;
            } else if (n == 4 && id_format != ANN_ID_INT64_BE) {
                uint32_t u = ((uint32_t)r.p[0] << 24) | ((uint32_t)r.p[1] << 16) | ((uint32_t)r.p[2] << 8) | r.p[3];
                *id = (int32_t)u;
            } else {
                char m[160];
                snprintf(m, sizeof(m), "BruteForceFileData: id of %d bytes; only Long (8, big-endian) and Int (4) injections are native "
                         "(AnnInjections.scala:8-12)", n);
                return perr(ANN_ERR_INVALID_ARGUMENT, m);
            }
            r.p += n;
            have_id = true;
        } else if (fid == 2 && type == T_STRUCT) {
            if (!walk_struct(r, &f, 0)) {
                // only running off the end of the file ends the stream quietly (TTransportException.END_OF_FILE,
                // ThriftIteratorIO.scala:42-49); a record that is malformed where it stands (negative size, unknown type,
                // absurd nesting) is an error there (TProtocolException propagates) and here
                if (f.bad || !r.eof) return perr(ANN_ERR_INVALID_ARGUMENT, "BruteForceFileData: malformed embedding struct");
                return 0;
            }
        } else if (!skip_value(r, type, 0)) {
            if (r.eof) return 0;
            return perr(ANN_ERR_INVALID_ARGUMENT, "BruteForceFileData: unknown thrift type in PersistedEmbedding");
        }
    }
    if (!have_id || !f.have) return perr(ANN_ERR_INVALID_ARGUMENT, "BruteForceFileData: PersistedEmbedding without id or embedding");
    return 1;
}

bool read_file(const std::string& path, std::vector<unsigned char>* out, std::string* why) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { *why = path + ": " + strerror(errno); return false; }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out->resize(n > 0 ? (size_t)n : 0);
    const bool ok = n <= 0 || fread(out->data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    if (!ok) *why = path + ": short read";
    return ok;
}

bool make_dir(const std::string& d) { return mkdir(d.c_str(), 0777) == 0 || errno == EEXIST; }

bool touch(const std::string& path) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fclose(f);
    return true;
}

int save_one(ann_index* ix, const std::string& dir, int id_format, int layout, int dim) {
    if (!make_dir(dir)) return perr(ANN_ERR_INVALID_ARGUMENT, "cannot create directory " + dir + ": " + strerror(errno));
    int64_t n = 0;
    int rc = ann_size(ix, &n);
    if (rc) return rc;
    Writer w;
    const std::string path = dir + "/" + kDataFile;
    w.f = fopen(path.c_str(), "wb");
    if (!w.f) return perr(ANN_ERR_INVALID_ARGUMENT, "cannot open " + path + ": " + strerror(errno));
    const int64_t chunk = 1 << 16;
    std::vector<int64_t> ids((size_t)std::min<int64_t>(chunk, std::max<int64_t>(n, 1)));
    std::vector<float> rows(ids.size() * (size_t)dim);
    for (int64_t s = 0; s < n; s += chunk) {      // insertion order, like linkedQueue.iterator() (BruteForceIndex.scala:152-155)
        const int64_t m = std::min(chunk, n - s);
        rc = ann_read_rows(ix, s, m, ids.data(), rows.data());
        if (rc) { fclose(w.f); return rc; }
        for (int64_t i = 0; i < m; ++i) {
            if (id_format == ANN_ID_INT32_BE && (ids[(size_t)i] < INT32_MIN || ids[(size_t)i] > INT32_MAX)) {
                fclose(w.f);
                return perr(ANN_ERR_INVALID_ARGUMENT, "id does not fit the Int injection");
            }
            write_record(w, ids[(size_t)i], id_format, rows.data() + (size_t)i * dim, dim, layout);
        }
        if (!w.flush()) { fclose(w.f); return perr(ANN_ERR_INVALID_ARGUMENT, "write to " + path + " failed"); }
    }
    if (fclose(w.f) != 0) return perr(ANN_ERR_INVALID_ARGUMENT, "closing " + path + " failed");
    return ANN_OK;
}

// stream one BruteForceFileData into `sink(ids, rows, n)` in batches; *dim is checked (or learnt when 0)
template <typename Sink>
int load_one(const std::string& dir, int id_format, int* dim, int64_t* total, Sink&& sink) {
    std::vector<unsigned char> bytes;
    std::string why;
    if (!read_file(dir + "/" + kDataFile, &bytes, &why)) return perr(ANN_ERR_INVALID_ARGUMENT, why);
    Reader r{bytes.data(), bytes.data() + bytes.size()};
    const size_t batch = 1 << 16;
    std::vector<int64_t> ids;
    std::vector<float> rows, one;
    ids.reserve(batch);
    for (;;) {
        int64_t id = 0;
        const int got = read_record(r, id_format, &id, &one);
        if (got < 0) return got;
        if (got == 1) {
            if (*dim == 0) *dim = (int)one.size();
            if ((int)one.size() != *dim) {
                char m[160];
                snprintf(m, sizeof(m), "BruteForceFileData: embedding of dimension %zu in an index of dimension %d", one.size(), *dim);
                return perr(ANN_ERR_DIMENSION_MISMATCH, m);
            }
            ids.push_back(id);
            rows.insert(rows.end(), one.begin(), one.end());
        }
        if (ids.size() == batch || (got == 0 && !ids.empty())) {
            int rc = sink(ids.data(), rows.data(), (int64_t)ids.size());
            if (rc) return rc;
            *total += (int64_t)ids.size();
            ids.clear();
            rows.clear();
        }
        if (got == 0) break;
    }
    return ANN_OK;
}

std::vector<std::string> shard_dirs(const std::string& dir) {
    std::vector<std::pair<long, std::string>> found;
    DIR* d = opendir(dir.c_str());
    if (!d) return {};
    while (dirent* e = readdir(d)) {
        const std::string name = e->d_name;
        if (name.rfind(kShardPrefix, 0) != 0) continue;
        char* endp = nullptr;
        const long i = strtol(name.c_str() + strlen(kShardPrefix), &endp, 10);
        if (endp && *endp == 0) found.emplace_back(i, dir + "/" + name);
    }
    closedir(d);
    std::sort(found.begin(), found.end());
    std::vector<std::string> out;
    for (auto& f : found) out.push_back(f.second);
    return out;
}

int check_formats(const char* who, int id_format, int layout) {
    if (id_format < 0 || id_format > ANN_ID_INT32_BE) return perr(ANN_ERR_INVALID_ARGUMENT, std::string(who) + ": unknown id_format");
    if (layout < 0 || layout > ANN_LAYOUT_RAW_FLOAT) return perr(ANN_ERR_INVALID_ARGUMENT, std::string(who) + ": unknown layout");
    return ANN_OK;
}

}  // namespace

extern "C" {

// The record codec by itself (pure host code, no device needed): what the byte-fixture tests drive.
int64_t ann_persisted_embedding_encode(int64_t id, int32_t id_format, const float* row, int32_t dim, int32_t layout,
                                       unsigned char* out, int64_t capacity) {
    if (!row || dim < 0 || check_formats("ann_persisted_embedding_encode", id_format, layout)) return -1;
    Writer w;
    write_record(w, id, id_format == ANN_ID_AUTO ? ANN_ID_INT64_BE : id_format, row, dim, layout);
    if (out && capacity >= (int64_t)w.buf.size()) memcpy(out, w.buf.data(), w.buf.size());
    return (int64_t)w.buf.size();
}

int ann_persisted_embedding_decode(const unsigned char* bytes, int64_t len, int32_t id_format, int64_t* id, float* row,
                                   int32_t row_capacity, int32_t* dim, int64_t* consumed) {
    if (!bytes || len < 0 || !id || !dim || !consumed) return perr(ANN_ERR_NULL_POINTER, "ann_persisted_embedding_decode: NULL argument");
    int rc = check_formats("ann_persisted_embedding_decode", id_format, 0);
    if (rc) return rc;
    Reader r{bytes, bytes + len};
    std::vector<float> one;
    const int got = read_record(r, id_format, id, &one);
    if (got < 0) return got;
    *consumed = got == 1 ? (int64_t)(r.p - bytes) : 0;   // 0 = end of stream (clean, or a truncated trailing record)
    *dim = got == 1 ? (int32_t)one.size() : 0;
    if (got == 1 && row && row_capacity >= (int32_t)one.size()) memcpy(row, one.data(), one.size() * sizeof(float));
    return ANN_OK;
}

int ann_save_directory(ann_index* ix, const char* directory, int32_t id_format, int32_t layout) {
    if (!ix || !directory) return perr(ANN_ERR_NULL_POINTER, "ann_save_directory: NULL argument");
    int rc = check_formats("ann_save_directory", id_format, layout);
    if (rc) return rc;
    int64_t dim = 0;
    rc = ann_get_stat(ix, "dim", &dim);
    if (rc) return rc;
    rc = save_one(ix, directory, id_format == ANN_ID_AUTO ? ANN_ID_INT64_BE : id_format, layout, (int)dim);
    if (rc) return rc;
    if (!touch(std::string(directory) + "/" + kSuccess)) return perr(ANN_ERR_INVALID_ARGUMENT, "cannot write the _SUCCESS marker");
    return ANN_OK;
}

int ann_load_directory(const ann_config* cfg, const char* directory, int32_t id_format, ann_index** out) {
    if (!cfg || !directory || !out) return perr(ANN_ERR_NULL_POINTER, "ann_load_directory: NULL argument");
    *out = nullptr;
    int rc = check_formats("ann_load_directory", id_format, 0);
    if (rc) return rc;
    ann_index* ix = nullptr;
    int dim = cfg->dim;
    int64_t total = 0;
    rc = load_one(directory, id_format, &dim, &total, [&](const int64_t* ids, const float* rows, int64_t n) -> int {
        if (!ix) {   // the dimension may come from the first record (cfg->dim == 0)
            ann_config c = *cfg;
            c.dim = dim;
            int r2 = ann_create(&c, &ix);
            if (r2) return r2;
        }
        return ann_append_batch(ix, ids, rows, n);
    });
    if (rc == ANN_OK && !ix) {   // an empty file is an empty index (needs the dimension from the caller)
        if (dim < 1) rc = perr(ANN_ERR_INVALID_ARGUMENT, "ann_load_directory: empty data file and cfg->dim == 0");
        else {
            ann_config c = *cfg;
            c.dim = dim;
            rc = ann_create(&c, &ix);
        }
    }
    if (rc) {
        ann_destroy(ix);
        return rc;
    }
    *out = ix;
    return ANN_OK;
}

int ann_sharded_save_directory(ann_sharded_index* sx, const char* directory, int32_t id_format, int32_t layout) {
    if (!sx || !directory) return perr(ANN_ERR_NULL_POINTER, "ann_sharded_save_directory: NULL argument");
    int rc = check_formats("ann_sharded_save_directory", id_format, layout);
    if (rc) return rc;
    if (!make_dir(directory)) return perr(ANN_ERR_INVALID_ARGUMENT, std::string("cannot create directory ") + directory);
    int64_t shards = 0, row_bytes = 0, n = 0;
    rc = ann_sharded_get_stat(sx, "shards", &shards);
    if (rc) return rc;
    for (int s = 0; s < (int)shards; ++s) {
        ann_index* ix = nullptr;
        int64_t rows = 0;
        rc = ann_sharded_shard(sx, s, &ix, &rows);
        if (rc) return rc;
        int64_t pitch_bytes = 0;
        rc = ann_get_stat(ix, "dim", &pitch_bytes);
        if (rc) return rc;
        rc = save_one(ix, std::string(directory) + "/" + kShardPrefix + std::to_string(s),
                      id_format == ANN_ID_AUTO ? ANN_ID_INT64_BE : id_format, layout, (int)pitch_bytes);
        if (rc) return rc;
    }
    (void)row_bytes;
    (void)n;
    if (!touch(std::string(directory) + "/" + kSuccess)) return perr(ANN_ERR_INVALID_ARGUMENT, "cannot write the _SUCCESS marker");
    return ANN_OK;
}

int ann_sharded_load_directory(const ann_config* cfg, const char* directory, int32_t id_format, const int32_t* device_ids,
                               int32_t n_devices, ann_sharded_index** out) {
    if (!cfg || !directory || !out) return perr(ANN_ERR_NULL_POINTER, "ann_sharded_load_directory: NULL argument");
    *out = nullptr;
    int rc = check_formats("ann_sharded_load_directory", id_format, 0);
    if (rc) return rc;
    std::vector<std::string> dirs = shard_dirs(directory);
    if (dirs.empty()) dirs.push_back(directory);    // an unsharded index directory loads into a sharded handle just as well
    ann_sharded_index* sx = nullptr;
    int dim = cfg->dim;
    int64_t total = 0;
    for (const std::string& d : dirs) {
        // rows are re-dealt over the devices batch by batch (the composed answer does not depend on which shard holds a row)
        rc = load_one(d, id_format, &dim, &total, [&](const int64_t* ids, const float* rows, int64_t n) -> int {
            if (!sx) {
                ann_config c = *cfg;
                c.dim = dim;
                int r2 = ann_sharded_create(&c, device_ids, n_devices, &sx);
                if (r2) return r2;
            }
            return ann_sharded_append_batch(sx, ids, rows, n);
        });
        if (rc) break;
    }
    if (rc == ANN_OK && !sx) {
        if (dim < 1) rc = perr(ANN_ERR_INVALID_ARGUMENT, "ann_sharded_load_directory: no rows and cfg->dim == 0");
        else {
            ann_config c = *cfg;
            c.dim = dim;
            rc = ann_sharded_create(&c, device_ids, n_devices, &sx);
        }
    }
    if (rc) {
        ann_sharded_destroy(sx);
        return rc;
    }
    *out = sx;
    return ANN_OK;
}

}  // extern "C"
