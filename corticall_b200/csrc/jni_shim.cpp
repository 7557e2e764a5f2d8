// jni_shim.cpp -- JNI bindings of libcorticall_cuda for java/uk/ac/ox/well/cortexjdk/utils/io/graph/cortex/NativeCortex.java.
// Argument marshalling ONLY: every function forwards to one cc_* entry point of include/corticall_cuda.h and throws
// uk.ac.ox.well.cortexjdk.utils.exceptions.CortexJDKException (S/utils/exceptions/CortexJDKException.java:3-10) with
// cc_last_error() on a non-zero status.
//
// Not built in the development image (no JDK, no jni.h): compiled only when <jni.h> is on the include path, e.g.
//   g++ -std=c++17 -O2 -fPIC -shared -I$JAVA_HOME/include -I$JAVA_HOME/include/linux jni_shim.cpp -L.. -lcorticall_cuda -o ../libcorticall_jni.so
#if __has_include(<jni.h>)
#include <jni.h>

#include <vector>

#include "../../include/corticall_cuda.h"

#define JFN(ret, name) extern "C" JNIEXPORT ret JNICALL Java_uk_ac_ox_well_cortexjdk_utils_io_graph_cortex_NativeCortex_##name

static bool check(JNIEnv *env, int status) {
    if (status == CC_OK) return true;
    jclass ex = env->FindClass("uk/ac/ox/well/cortexjdk/utils/exceptions/CortexJDKException");
    if (ex) env->ThrowNew(ex, cc_last_error());
    return false;
}
static cc_graph *G(jlong h) { return reinterpret_cast<cc_graph *>(h); }
static cc_sharded *SH(jlong h) { return reinterpret_cast<cc_sharded *>(h); }

static bool fail_arg(JNIEnv *env, const char *what) {
    jclass ex = env->FindClass("uk/ac/ox/well/cortexjdk/utils/exceptions/CortexJDKException");
    if (ex) env->ThrowNew(ex, what);
    return false;
}
// A Java array must hold at least `need` elements before native code writes `need` elements into it (or reads them).
static bool has(JNIEnv *env, jarray a, uint64_t need, const char *what) {
    if (!a) return need == 0 ? true : fail_arg(env, what);
    return (uint64_t)env->GetArrayLength(a) >= need ? true : fail_arg(env, what);
}
static void graph_shape(jlong h, uint32_t *k, uint32_t *s, uint32_t *c, uint64_t *S) {
    cc_header(G(h), nullptr, k, s, c, nullptr, nullptr, S);
}

JFN(jlong, open)(JNIEnv *env, jclass, jstring path, jint device) {
    const char *p = path ? env->GetStringUTFChars(path, nullptr) : nullptr;
    if (!p) { fail_arg(env, "null path"); return 0; }
    cc_graph *g = nullptr;
    const int rc = cc_open(p, device, &g);
    env->ReleaseStringUTFChars(path, p);
    return check(env, rc) ? reinterpret_cast<jlong>(g) : 0;
}
JFN(void, dispose)(JNIEnv *, jclass, jlong h) { cc_dispose(G(h)); }

JFN(jlongArray, header)(JNIEnv *env, jclass, jlong h) {
    uint32_t v, k, s, c; uint64_t n, off, rs;
    if (!check(env, cc_header(G(h), &v, &k, &s, &c, &n, &off, &rs))) return nullptr;
    const jlong vals[7] = {(jlong)v, (jlong)k, (jlong)s, (jlong)c, (jlong)n, (jlong)off, (jlong)rs};
    jlongArray out = env->NewLongArray(7);
    env->SetLongArrayRegion(out, 0, 7, vals);
    return out;
}
JFN(jstring, colorName)(JNIEnv *env, jclass, jlong h, jint color) {
    std::vector<char> buf(1 << 16);
    return check(env, cc_color_name(G(h), (uint32_t)color, buf.data(), buf.size())) ? env->NewStringUTF(buf.data()) : nullptr;
}
JFN(jstring, colorGraphName)(JNIEnv *env, jclass, jlong h, jint color) {
    std::vector<char> buf(1 << 16);
    return check(env, cc_color_graph_name(G(h), (uint32_t)color, buf.data(), buf.size())) ? env->NewStringUTF(buf.data()) : nullptr;
}
JFN(jlongArray, colorInfo)(JNIEnv *env, jclass, jlong h, jint color) {
    cc_color_info ci;
    if (!check(env, cc_color_info_get(G(h), (uint32_t)color, &ci))) return nullptr;
    const jlong vals[8] = {(jlong)ci.mean_read_length, (jlong)ci.total_sequence, ci.tip_clipping, ci.low_covg_supernodes_removed,
                           ci.low_covg_kmers_removed, ci.cleaned_against_graph, (jlong)ci.low_cov_supernodes_threshold,
                           (jlong)ci.low_cov_kmer_threshold};
    jlongArray out = env->NewLongArray(8);
    env->SetLongArrayRegion(out, 0, 8, vals);
    return out;
}

JFN(void, decodeRecords)(JNIEnv *env, jclass, jlong h, jlong first, jint count, jlongArray kmers, jintArray cov, jbyteArray edges) {
    uint32_t s = 0, nc = 0;
    graph_shape(h, nullptr, &s, &nc, nullptr);
    if (count < 0 || !has(env, kmers, (uint64_t)count * s, "binaryKmers shorter than count * kmerBits") ||
        !has(env, cov, (uint64_t)count * nc, "coverages shorter than count * numColors") ||
        !has(env, edges, (uint64_t)count * nc, "edges shorter than count * numColors")) return;
    jlong *k = env->GetLongArrayElements(kmers, nullptr);
    jint *c = env->GetIntArrayElements(cov, nullptr);
    jbyte *e = env->GetByteArrayElements(edges, nullptr);
    if (!k || !c || !e) {
        if (k) env->ReleaseLongArrayElements(kmers, k, JNI_ABORT);
        if (c) env->ReleaseIntArrayElements(cov, c, JNI_ABORT);
        if (e) env->ReleaseByteArrayElements(edges, e, JNI_ABORT);
        fail_arg(env, "out of memory pinning arrays");
        return;
    }
    const int rc = cc_decode_records(G(h), (uint64_t)first, (uint64_t)count, reinterpret_cast<uint64_t *>(k),
                                     reinterpret_cast<int32_t *>(c), reinterpret_cast<uint8_t *>(e));
    if (rc == CC_OK) {     // Java long[] = Long.reverseBytes(native word) (CortexGraph.java:208-209): only the words just written
        for (uint64_t i = 0; i < (uint64_t)count * s; ++i) k[i] = (jlong)__builtin_bswap64((uint64_t)k[i]);
    }
    env->ReleaseLongArrayElements(kmers, k, 0);
    env->ReleaseIntArrayElements(cov, c, 0);
    env->ReleaseByteArrayElements(edges, e, 0);
    check(env, rc);
}

JFN(jlong, findNovel)(JNIEnv *env, jclass, jlong h, jint child, jintArray parents, jbyteArray outRecords, jlongArray outIndex) {
    uint32_t s = 0;
    cc_header(G(h), nullptr, nullptr, &s, nullptr, nullptr, nullptr, nullptr);
    const jsize np = env->GetArrayLength(parents);
    jint *p = env->GetIntArrayElements(parents, nullptr);
    jbyte *r = outRecords ? env->GetByteArrayElements(outRecords, nullptr) : nullptr;
    jlong *ix = outIndex ? env->GetLongArrayElements(outIndex, nullptr) : nullptr;
    uint64_t cap = r ? (uint64_t)env->GetArrayLength(outRecords) / (8ull * s + 5) : 0, total = 0;
    if (ix) cap = cap < (uint64_t)env->GetArrayLength(outIndex) ? cap : (uint64_t)env->GetArrayLength(outIndex);
    const int rc = cc_find_novel(G(h), child, reinterpret_cast<int32_t *>(p), np, r, reinterpret_cast<uint64_t *>(ix), cap, &total);
    env->ReleaseIntArrayElements(parents, p, JNI_ABORT);
    if (r) env->ReleaseByteArrayElements(outRecords, r, 0);
    if (ix) env->ReleaseLongArrayElements(outIndex, ix, 0);
    return check(env, rc) ? (jlong)total : -1;
}
JFN(jlong, writeRoiFile)(JNIEnv *env, jclass, jlong h, jint child, jintArray parents, jstring outPath) {
    const jsize np = env->GetArrayLength(parents);
    jint *p = env->GetIntArrayElements(parents, nullptr);
    const char *path = env->GetStringUTFChars(outPath, nullptr);
    uint64_t total = 0;
    const int rc = cc_write_roi_file(G(h), child, reinterpret_cast<int32_t *>(p), np, path, &total);
    env->ReleaseStringUTFChars(outPath, path);
    env->ReleaseIntArrayElements(parents, p, JNI_ABORT);
    return check(env, rc) ? (jlong)total : -1;
}

JFN(void, findAscii)(JNIEnv *env, jclass, jlong h, jbyteArray kmers, jint nq, jlongArray outIndex) {
    uint32_t kk = 0;
    graph_shape(h, &kk, nullptr, nullptr, nullptr);
    if (nq < 0 || !has(env, kmers, (uint64_t)nq * kk, "kmers shorter than nq * kmerSize") || !has(env, outIndex, (uint64_t)nq, "outIndex shorter than nq")) return;
    jbyte *q = env->GetByteArrayElements(kmers, nullptr);
    jlong *o = env->GetLongArrayElements(outIndex, nullptr);
    const int rc = cc_find_ascii(G(h), reinterpret_cast<uint8_t *>(q), (uint64_t)nq, reinterpret_cast<int64_t *>(o), CC_ALGO_AUTO);
    env->ReleaseByteArrayElements(kmers, q, JNI_ABORT);
    env->ReleaseLongArrayElements(outIndex, o, 0);
    check(env, rc);
}
JFN(void, findWindows)(JNIEnv *env, jclass, jlong h, jbyteArray seq, jlongArray outIndex) {
    uint32_t kk = 0;
    graph_shape(h, &kk, nullptr, nullptr, nullptr);
    const uint64_t len = seq ? (uint64_t)env->GetArrayLength(seq) : 0;
    if (!seq || !has(env, outIndex, len >= kk ? len - kk + 1 : 0, "outIndex shorter than the number of windows")) { if (!seq) fail_arg(env, "null sequence"); return; }
    jbyte *q = env->GetByteArrayElements(seq, nullptr);
    jlong *o = env->GetLongArrayElements(outIndex, nullptr);
    const int rc = cc_find_windows(G(h), reinterpret_cast<uint8_t *>(q), (uint64_t)env->GetArrayLength(seq),
                                   reinterpret_cast<int64_t *>(o), CC_ALGO_AUTO);
    env->ReleaseByteArrayElements(seq, q, JNI_ABORT);
    env->ReleaseLongArrayElements(outIndex, o, 0);
    check(env, rc);
}
JFN(void, containsWindows)(JNIEnv *env, jclass, jlong h, jbyteArray seq, jbooleanArray outPresent) {
    uint32_t kk = 0;
    graph_shape(h, &kk, nullptr, nullptr, nullptr);
    const uint64_t len = seq ? (uint64_t)env->GetArrayLength(seq) : 0;
    if (!seq || !has(env, outPresent, len >= kk ? len - kk + 1 : 0, "outPresent shorter than the number of windows")) { if (!seq) fail_arg(env, "null sequence"); return; }
    jbyte *q = env->GetByteArrayElements(seq, nullptr);
    jboolean *o = env->GetBooleanArrayElements(outPresent, nullptr);
    const int rc = cc_contains_windows(G(h), reinterpret_cast<uint8_t *>(q), (uint64_t)env->GetArrayLength(seq),
                                       reinterpret_cast<uint8_t *>(o));
    env->ReleaseByteArrayElements(seq, q, JNI_ABORT);
    env->ReleaseBooleanArrayElements(outPresent, o, 0);
    check(env, rc);
}
JFN(void, packCanonical)(JNIEnv *env, jclass, jint device, jbyteArray seq, jint k, jlongArray outKmers, jbyteArray outFlags) {
    const uint64_t len = seq ? (uint64_t)env->GetArrayLength(seq) : 0;
    const uint64_t nwin = (k > 0 && len >= (uint64_t)k) ? len - (uint64_t)k + 1 : 0, sw = k > 0 ? ((uint64_t)k + 31) / 32 : 0;
    if (!seq || k <= 0 || !has(env, outKmers, nwin * sw, "outBinaryKmers shorter than windows * kmerBits") ||
        !has(env, outFlags, nwin, "outFlags shorter than the number of windows")) { if (!seq || k <= 0) fail_arg(env, "bad sequence or k"); return; }
    jbyte *q = env->GetByteArrayElements(seq, nullptr);
    jlong *w = env->GetLongArrayElements(outKmers, nullptr);
    jbyte *f = env->GetByteArrayElements(outFlags, nullptr);
    const int rc = cc_pack_canonical(device, reinterpret_cast<uint8_t *>(q), len, (uint32_t)k,
                                     reinterpret_cast<uint64_t *>(w), reinterpret_cast<uint8_t *>(f));
    if (rc == CC_OK) {     // only the words just written
        for (uint64_t i = 0; i < nwin * sw; ++i) w[i] = (jlong)__builtin_bswap64((uint64_t)w[i]);
    }
    env->ReleaseByteArrayElements(seq, q, JNI_ABORT);
    env->ReleaseLongArrayElements(outKmers, w, 0);
    env->ReleaseByteArrayElements(outFlags, f, 0);
    check(env, rc);
}

// ---- next rows: whole-graph operations returning a new device-resident graph
JFN(jlong, join)(JNIEnv *env, jclass, jlongArray handles) {
    const jsize n = env->GetArrayLength(handles);
    jlong *h = env->GetLongArrayElements(handles, nullptr);
    std::vector<cc_graph *> gs(n);
    for (jsize i = 0; i < n; ++i) gs[i] = G(h[i]);
    env->ReleaseLongArrayElements(handles, h, JNI_ABORT);
    cc_graph *out = nullptr;
    return check(env, cc_join(gs.data(), (int)n, &out)) ? reinterpret_cast<jlong>(out) : 0;
}
JFN(jlong, sort)(JNIEnv *env, jclass, jlong h) {
    cc_graph *out = nullptr;
    return check(env, cc_sort(G(h), &out)) ? reinterpret_cast<jlong>(out) : 0;
}
JFN(void, writeGraph)(JNIEnv *env, jclass, jlong h, jstring outPath) {
    const char *path = env->GetStringUTFChars(outPath, nullptr);
    const int rc = cc_write_graph(G(h), path);
    env->ReleaseStringUTFChars(outPath, path);
    check(env, rc);
}
JFN(jlong, findLowCoverage)(JNIEnv *env, jclass, jlong roi, jint minCoverage) {
    cc_graph *out = nullptr;
    return check(env, cc_find_low_coverage(G(roi), minCoverage, &out)) ? reinterpret_cast<jlong>(out) : 0;
}
JFN(jlong, findShared)(JNIEnv *env, jclass, jlong graph, jlong roi, jint child, jintArray parents, jintArray ignore) {
    const jsize np = env->GetArrayLength(parents), ni = env->GetArrayLength(ignore);
    jint *p = env->GetIntArrayElements(parents, nullptr);
    jint *g = env->GetIntArrayElements(ignore, nullptr);
    cc_graph *out = nullptr;
    const int rc = cc_find_shared(G(graph), G(roi), child, reinterpret_cast<int32_t *>(p), np, reinterpret_cast<int32_t *>(g), ni, &out);
    env->ReleaseIntArrayElements(parents, p, JNI_ABORT);
    env->ReleaseIntArrayElements(ignore, g, JNI_ABORT);
    return check(env, rc) ? reinterpret_cast<jlong>(out) : 0;
}
JFN(jlong, recoverExcludedKmers)(JNIEnv *env, jclass, jlong graph, jlong dirty, jint child) {
    cc_graph *out = nullptr;
    uint64_t recovered = 0;
    return check(env, cc_recover_excluded_kmers(G(graph), G(dirty), child, &out, &recovered)) ? reinterpret_cast<jlong>(out) : 0;
}
JFN(jintArray, covStats)(JNIEnv *env, jclass, jlong h, jint child, jintArray parents) {
    const jsize np = env->GetArrayLength(parents);
    jint *p = env->GetIntArrayElements(parents, nullptr);
    uint64_t rows = 0;
    int rc = cc_cov_stats(G(h), child, reinterpret_cast<int32_t *>(p), np, nullptr, nullptr, 0, &rows);
    std::vector<int32_t> cov(rows ? rows : 1), cnt(rows ? rows : 1);
    if (rc == CC_OK) rc = cc_cov_stats(G(h), child, reinterpret_cast<int32_t *>(p), np, cov.data(), cnt.data(), rows, &rows);
    env->ReleaseIntArrayElements(parents, p, JNI_ABORT);
    if (!check(env, rc)) return nullptr;
    std::vector<jint> flat(2 * rows);
    for (uint64_t i = 0; i < rows; ++i) { flat[2 * i] = cov[i]; flat[2 * i + 1] = cnt[i]; }
    jintArray out = env->NewIntArray((jsize)flat.size());
    env->SetIntArrayRegion(out, 0, (jsize)flat.size(), flat.data());
    return out;
}

// ---- the legacy per-record findRecord, batched: one native call and one kernel launch for a vertex and its neighbours
JFN(void, findRecords)(JNIEnv *env, jclass, jlong h, jbyteArray kmers, jint nq, jlongArray outIndex, jlongArray binaryKmers, jintArray coverages,
                       jbyteArray edges) {
    uint32_t kk = 0, s = 0, nc = 0; uint64_t S = 0;
    graph_shape(h, &kk, &s, &nc, &S);
    if (nq < 0 || !has(env, kmers, (uint64_t)nq * kk, "kmers shorter than nq * kmerSize") || !has(env, outIndex, (uint64_t)nq, "outIndex shorter than nq") ||
        !has(env, binaryKmers, (uint64_t)nq * s, "binaryKmers shorter than nq * kmerBits") || !has(env, coverages, (uint64_t)nq * nc, "coverages shorter than nq * numColors") ||
        !has(env, edges, (uint64_t)nq * nc, "edges shorter than nq * numColors")) return;
    jbyte *q = env->GetByteArrayElements(kmers, nullptr);
    std::vector<int64_t> idx((size_t)nq ? (size_t)nq : 1);
    std::vector<uint8_t> raw((size_t)nq * S + 1);
    const int rc = q ? cc_find_records(G(h), reinterpret_cast<uint8_t *>(q), (uint64_t)nq, idx.data(), raw.data()) : CC_ERR_ARG;
    if (q) env->ReleaseByteArrayElements(kmers, q, JNI_ABORT);
    if (!check(env, rc)) return;
    std::vector<jlong> bk((size_t)nq * s + 1);
    std::vector<jint> cv((size_t)nq * nc + 1);
    std::vector<jbyte> ed((size_t)nq * nc + 1);
    for (jint i = 0; i < nq; ++i) {
        const uint8_t *r = raw.data() + (size_t)i * S;
        for (uint32_t w = 0; w < s; ++w) {            // on-disk word read big-endian = the Java long (CortexGraph.java:208-209)
            uint64_t v = 0;
            for (int b = 0; b < 8; ++b) v = (v << 8) | r[8 * w + b];
            bk[(size_t)i * s + w] = (jlong)v;
        }
        for (uint32_t c = 0; c < nc; ++c) {
            const uint8_t *p = r + 8 * s + 4 * c;
            cv[(size_t)i * nc + c] = (jint)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
            ed[(size_t)i * nc + c] = (jbyte)r[8 * s + 4 * nc + c];
        }
    }
    env->SetLongArrayRegion(outIndex, 0, nq, reinterpret_cast<const jlong *>(idx.data()));
    env->SetLongArrayRegion(binaryKmers, 0, (jsize)(nq * s), bk.data());
    env->SetIntArrayRegion(coverages, 0, (jsize)(nq * nc), cv.data());
    env->SetByteArrayRegion(edges, 0, (jsize)(nq * nc), ed.data());
}

JFN(jlongArray, remove)(JNIEnv *env, jclass, jlong primary, jlongArray secondaries) {
    const jsize n = secondaries ? env->GetArrayLength(secondaries) : 0;
    std::vector<cc_graph *> gs((size_t)n + 1);
    if (n) {
        jlong *h = env->GetLongArrayElements(secondaries, nullptr);
        if (!h) { fail_arg(env, "out of memory pinning arrays"); return nullptr; }
        for (jsize i = 0; i < n; ++i) gs[i] = G(h[i]);
        env->ReleaseLongArrayElements(secondaries, h, JNI_ABORT);
    }
    cc_graph *out = nullptr;
    uint64_t removed = 0;
    if (!check(env, cc_remove(G(primary), gs.data(), (int)n, &out, &removed))) return nullptr;
    const jlong vals[2] = {reinterpret_cast<jlong>(out), (jlong)removed};
    jlongArray res = env->NewLongArray(2);
    env->SetLongArrayRegion(res, 0, 2, vals);
    return res;
}

// ---- one graph over several GPUs of this process
JFN(jlong, openSharded)(JNIEnv *env, jclass, jstring path, jintArray devices, jint placement) {
    const char *p = path ? env->GetStringUTFChars(path, nullptr) : nullptr;
    if (!p || !devices) { if (p) env->ReleaseStringUTFChars(path, p); fail_arg(env, "null path or device list"); return 0; }
    const jsize n = env->GetArrayLength(devices);
    jint *d = env->GetIntArrayElements(devices, nullptr);
    cc_sharded *sh = nullptr;
    const int rc = d ? cc_open_sharded_placed(p, reinterpret_cast<int *>(d), (int)n, (int)placement, &sh) : CC_ERR_ARG;
    if (d) env->ReleaseIntArrayElements(devices, d, JNI_ABORT);
    env->ReleaseStringUTFChars(path, p);
    return check(env, rc) ? reinterpret_cast<jlong>(sh) : 0;
}
JFN(void, disposeSharded)(JNIEnv *, jclass, jlong h) { cc_dispose_sharded(SH(h)); }
JFN(jlongArray, shardedInfo)(JNIEnv *env, jclass, jlong h) {
    int nd = 0; uint64_t n = 0; uint32_t k = 0, c = 0;
    if (!check(env, cc_sharded_info(SH(h), &nd, &n, &k, &c))) return nullptr;
    const jlong vals[4] = {(jlong)nd, (jlong)n, (jlong)k, (jlong)c};
    jlongArray out = env->NewLongArray(4);
    env->SetLongArrayRegion(out, 0, 4, vals);
    return out;
}
JFN(jlongArray, shardedShard)(JNIEnv *env, jclass, jlong h, jint rank) {
    cc_graph *g = nullptr; int dev = 0; uint64_t first = 0;
    if (!check(env, cc_sharded_shard(SH(h), rank, &g, &dev, &first))) return nullptr;
    const jlong vals[3] = {reinterpret_cast<jlong>(g), (jlong)dev, (jlong)first};
    jlongArray out = env->NewLongArray(3);
    env->SetLongArrayRegion(out, 0, 3, vals);
    return out;
}
JFN(void, findAsciiSharded)(JNIEnv *env, jclass, jlong h, jbyteArray kmers, jint nq, jlongArray outIndex) {
    uint32_t kk = 0;
    if (!check(env, cc_sharded_info(SH(h), nullptr, nullptr, &kk, nullptr))) return;
    if (nq < 0 || !has(env, kmers, (uint64_t)nq * kk, "kmers shorter than nq * kmerSize") || !has(env, outIndex, (uint64_t)nq, "outIndex shorter than nq")) return;
    jbyte *q = env->GetByteArrayElements(kmers, nullptr);
    jlong *o = env->GetLongArrayElements(outIndex, nullptr);
    const int rc = (q && o) ? cc_find_ascii_sharded(SH(h), reinterpret_cast<uint8_t *>(q), (uint64_t)nq, reinterpret_cast<int64_t *>(o)) : CC_ERR_ARG;
    if (q) env->ReleaseByteArrayElements(kmers, q, JNI_ABORT);
    if (o) env->ReleaseLongArrayElements(outIndex, o, 0);
    check(env, rc);
}
JFN(void, findWindowsSharded)(JNIEnv *env, jclass, jlong h, jbyteArray seq, jlongArray outIndex) {
    uint32_t kk = 0;
    if (!check(env, cc_sharded_info(SH(h), nullptr, nullptr, &kk, nullptr))) return;
    const uint64_t len = seq ? (uint64_t)env->GetArrayLength(seq) : 0;
    if (!seq || !has(env, outIndex, len >= kk ? len - kk + 1 : 0, "outIndex shorter than the number of windows")) { if (!seq) fail_arg(env, "null sequence"); return; }
    jbyte *q = env->GetByteArrayElements(seq, nullptr);
    jlong *o = env->GetLongArrayElements(outIndex, nullptr);
    const int rc = (q && o) ? cc_find_windows_sharded(SH(h), reinterpret_cast<uint8_t *>(q), len, reinterpret_cast<int64_t *>(o)) : CC_ERR_ARG;
    if (q) env->ReleaseByteArrayElements(seq, q, JNI_ABORT);
    if (o) env->ReleaseLongArrayElements(outIndex, o, 0);
    check(env, rc);
}
JFN(jlong, findNovelSharded)(JNIEnv *env, jclass, jlong h, jint child, jintArray parents, jbyteArray outRecords, jlongArray outIndex) {
    uint32_t kk = 0;
    if (!check(env, cc_sharded_info(SH(h), nullptr, nullptr, &kk, nullptr)) || !parents) { if (!parents) fail_arg(env, "null parent list"); return -1; }
    const uint64_t O = 8ull * ((kk + 31) / 32) + 5;
    const jsize np = env->GetArrayLength(parents);
    jint *p = env->GetIntArrayElements(parents, nullptr);
    jbyte *r = outRecords ? env->GetByteArrayElements(outRecords, nullptr) : nullptr;
    jlong *ix = outIndex ? env->GetLongArrayElements(outIndex, nullptr) : nullptr;
    uint64_t cap = r ? (uint64_t)env->GetArrayLength(outRecords) / O : 0, total = 0;
    if (ix) cap = cap < (uint64_t)env->GetArrayLength(outIndex) ? cap : (uint64_t)env->GetArrayLength(outIndex);
    const int rc = p ? cc_find_novel_sharded(SH(h), child, reinterpret_cast<int32_t *>(p), np, r, reinterpret_cast<uint64_t *>(ix), cap, &total) : CC_ERR_ARG;
    if (p) env->ReleaseIntArrayElements(parents, p, JNI_ABORT);
    if (r) env->ReleaseByteArrayElements(outRecords, r, 0);
    if (ix) env->ReleaseLongArrayElements(outIndex, ix, 0);
    return check(env, rc) ? (jlong)total : -1;
}
JFN(jlong, writeRoiFileSharded)(JNIEnv *env, jclass, jlong h, jint child, jintArray parents, jstring outPath) {
    if (!parents || !outPath) { fail_arg(env, "null argument"); return -1; }
    const jsize np = env->GetArrayLength(parents);
    jint *p = env->GetIntArrayElements(parents, nullptr);
    const char *path = env->GetStringUTFChars(outPath, nullptr);
    uint64_t total = 0;
    const int rc = (p && path) ? cc_write_roi_file_sharded(SH(h), child, reinterpret_cast<int32_t *>(p), np, path, &total) : CC_ERR_ARG;
    if (path) env->ReleaseStringUTFChars(outPath, path);
    if (p) env->ReleaseIntArrayElements(parents, p, JNI_ABORT);
    return check(env, rc) ? (jlong)total : -1;
}
#endif  // __has_include(<jni.h>)
