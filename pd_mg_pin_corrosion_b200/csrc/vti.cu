// vti.cu -- device-side formatting of the reference's VTI snapshot (src/vtk_writer.cpp:16-146).
//
// The reference writes every output interval (src/coupling.cpp:242-249) an ASCII ImageData file:
// one line per node and DataArray, doubles through `ostream << double` (= printf "%g", six
// significant digits), WALL/OUTSIDE velocities zeroed, NaN/Inf and |v| < 1e-300 flushed to 0.
// On the CPU that is ~4 us per node; here the text is produced on the device:
//   k_rec_len    one thread per node formats its record and keeps only the length
//   scan         exclusive prefix sum of the lengths (scan.cuh)
//   k_rec_write  a block formats its records into shared memory at their final relative positions
//                and writes its piece of the DataArray body with aligned 16-byte stores
// and only the finished text crosses PCIe. Output is byte-identical to the reference's file.
//
// "%g" needs the correctly rounded 6-digit decimal of a binary64 value. x = m 2^e is multiplied
// by a 128-bit power of ten from pow10_table.inc (T 2^b <= 10^k < (T+1) 2^b): the 181-bit product
// brackets x 10^k within m 2^(e+b), 2^-107 relative. A 53-bit value cannot come that close to a
// 7-digit half-way point without being exactly on it, which happens only for integers
// (x = (2N+1) 5^j 2^(j-1)) and is decided by an exact 128-bit comparison (round half to even, as
// glibc). Anything still undecided raises the error flag of the call (never observed).
#include <algorithm>
#include <sstream>

#include "common.cuh"
#include "iopool.cuh"
#include "scan.cuh"

struct Pow10Entry {
    unsigned long long hi, lo;
    int bexp, exact;
};
#include "pow10_table.inc"

namespace {

__constant__ Pow10Entry d_pow10[PD_POW10_KMAX - PD_POW10_KMIN + 1];
bool g_table_uploaded[64] = {false};

typedef unsigned __int128 u128;

// digits of |v| (normal, finite, non-zero): q in [100000, 999999], decimal exponent X with
// |v| ~= q * 10^(X-5) correctly rounded (half to even). *flag is set if the decision was not provable.
__device__ __forceinline__ void dec6(double av, unsigned* q_out, int* X_out, int* flag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(av);
    const int be = (int)((bits >> 52) & 0x7ff);
    const unsigned long long m = (bits & 0xfffffffffffffull) | (1ull << 52);   // 2^52 <= m < 2^53
    const int e = be - 1075;                                                    // av = m 2^e
    const int e2 = be - 1023;                                                   // floor(log2 av)
    int X = (e2 * 78913) >> 18;                                                 // floor(e2 log10 2) <= floor(log10 av) <= that + 1
    unsigned long long I = 0;
    bool up = false;
    for (int attempt = 0; attempt < 3; ++attempt) {
        const int k = 5 - X;
        const Pow10Entry T = d_pow10[k - PD_POW10_KMIN];
        // P = m * (T.hi:T.lo), 192 bits p2:p1:p0
        const unsigned long long p0 = m * T.lo;
        const unsigned long long c0 = __umul64hi(m, T.lo);
        const unsigned long long t1 = m * T.hi;
        const unsigned long long p1 = t1 + c0;
        const unsigned long long p2 = __umul64hi(m, T.hi) + (p1 < t1 ? 1ull : 0ull);
        const int s = -(e + T.bexp);                  // value = P 2^-s, 128 < s < 192
        const int sh = s - 128;                       // integer part = p2 >> sh
        I = p2 >> sh;
        if (I >= 1000000ull) { ++X; continue; }       // one decimal digit more than estimated
        if (I < 99999ull) { --X; continue; }          // (estimate one too high: negative binary exponents)
        // fraction F = (p2 & mask):p1:p0 against half = 2^(s-1)
        const unsigned long long fmask = (1ull << sh) - 1ull;
        const unsigned long long f2 = p2 & fmask;
        const unsigned long long halfbit = 1ull << (sh - 1);
        const bool ge_half = (f2 & halfbit) != 0;
        const bool rest_zero = ((f2 & (halfbit - 1ull)) | p1 | p0) == 0ull;
        if (T.exact) {
            if (ge_half && rest_zero) up = (I & 1ull) != 0;          // exact tie: half to even
            else up = ge_half;
        } else if (ge_half) {
            up = true;                                                // 10^k > T 2^b strictly
        } else {
            // upper bound P + m (exclusive): still below half?
            const unsigned long long q0 = p0 + m;
            const unsigned long long cq = q0 < p0 ? 1ull : 0ull;
            const unsigned long long q1 = p1 + cq;
            const unsigned long long q2 = p2 + ((cq && q1 == 0ull) ? 1ull : 0ull);
            const bool same_int = (q2 >> sh) == I;
            const unsigned long long g2 = q2 & fmask;
            const bool up_ge_half = !same_int || (g2 & halfbit) != 0;
            const bool up_is_half = same_int && (g2 & halfbit) != 0 && ((g2 & (halfbit - 1ull)) | q1 | q0) == 0ull;
            if (!up_ge_half || up_is_half) {
                up = false;                                           // x 10^k < upper <= half
            } else {
                // the bracket straddles the half-way point: only an exact tie can do that,
                // m 2^(e+1) == (2I+1) 10^-k with -k in 1..22
                bool tie = false;
                if (k < 0 && -k <= 22) {
                    u128 R = (u128)(2ull * I + 1ull);
                    for (int t = 0; t < -k; ++t) R *= 10u;
                    const int sl = e + 1;
                    if (sl >= 0) {
                        if (sl <= 70) tie = (((u128)m) << sl) == R;
                    } else if (-sl < 64) {
                        tie = ((m & ((1ull << (-sl)) - 1ull)) == 0ull) && ((u128)(m >> (-sl)) == R);
                    }
                }
                if (tie) up = (I & 1ull) != 0;
                else { up = true; *flag = 1; }
            }
        }
        // I = 99999 is a 6-digit result only if it rounds up to 100000 (x = 10^X seen through a truncated 10^k)
        if (I + (up ? 1ull : 0ull) < 100000ull) { --X; continue; }
        break;
    }
    unsigned q = (unsigned)I + (up ? 1u : 0u);
    if (q >= 1000000u) { q = 100000u; ++X; }
    *q_out = q;
    *X_out = X;
}

// printf("%g", safe_val(v)) into out (<= 13 chars); returns the length
__device__ __forceinline__ int fmt_g(double v, char* out, int* flag) {
    // safe_val (src/vtk_writer.cpp:8-14)
    if (isnan(v) || isinf(v)) v = 0.0;
    if (v != 0.0 && fabs(v) < 1e-300) v = 0.0;
    int n = 0;
    if (__double_as_longlong(v) < 0) out[n++] = '-';
    if (v == 0.0) { out[n++] = '0'; return n; }
    unsigned q;
    int X;
    dec6(fabs(v), &q, &X, flag);
    char d[6];
#pragma unroll
    for (int t = 5; t >= 0; --t) { d[t] = (char)('0' + q % 10u); q /= 10u; }
    int nd = 6;
    while (nd > 1 && d[nd - 1] == '0') --nd;          // %g strips trailing zeros
    if (X < -4 || X >= 6) {                           // exponential style
        out[n++] = d[0];
        if (nd > 1) {
            out[n++] = '.';
            for (int t = 1; t < nd; ++t) out[n++] = d[t];
        }
        out[n++] = 'e';
        int ax = X;
        if (X < 0) { out[n++] = '-'; ax = -X; } else out[n++] = '+';
        if (ax >= 100) { out[n++] = (char)('0' + ax / 100); ax %= 100; out[n++] = (char)('0' + ax / 10); }
        else out[n++] = (char)('0' + ax / 10);
        out[n++] = (char)('0' + ax % 10);
    } else if (X >= 0) {                              // fixed, integer part of X+1 digits
        for (int t = 0; t <= X; ++t) out[n++] = d[t];
        if (nd > X + 1) {
            out[n++] = '.';
            for (int t = X + 1; t < nd; ++t) out[n++] = d[t];
        }
    } else {                                          // 0.000ddd
        out[n++] = '0';
        out[n++] = '.';
        for (int t = 0; t < -X - 1; ++t) out[n++] = '0';
        for (int t = 0; t < nd; ++t) out[n++] = d[t];
    }
    return n;
}

__device__ __forceinline__ int fmt_int(int v, char* out) {
    int n = 0;
    unsigned u = (unsigned)v;
    if (v < 0) { out[n++] = '-'; u = (unsigned)(-(long long)v); }
    char tmp[10];
    int t = 0;
    do { tmp[t++] = (char)('0' + u % 10u); u /= 10u; } while (u);
    while (t) out[n++] = tmp[--t];
    return n;
}

constexpr int kIndent = 10;    // "          " in front of every record
constexpr int kMaxRec = 56;    // longest record: indent + 3 x 13 chars + 2 blanks + newline = 52

__device__ __forceinline__ int put_indent(char* o) {
#pragma unroll
    for (int t = 0; t < kIndent; ++t) o[t] = ' ';
    return kIndent;
}

enum RecKind { R_VEL2, R_VEL3, R_F64, R_U8, R_I32 };
struct RecArgs {
    const double *a, *b, *c;       // scalar field or velocity components (owned node 0)
    const uint8_t* type;           // node types (velocity: WALL/OUTSIDE print 0, src/vtk_writer.cpp:62)
    const void* ints;              // uint8 / int32 arrays
};

// one record (= one line of a DataArray body) of node t into o; returns its length
template <int KIND>
__device__ __forceinline__ int make_record(const RecArgs& A, long long t, char* o, int* fl) {
    int k = put_indent(o);
    if (KIND == R_VEL2 || KIND == R_VEL3) {
        const uint8_t ty = A.type[t];
        const bool fict = (ty == PDGPU_WALL || ty == PDGPU_OUTSIDE);
        k += fmt_g(fict ? 0.0 : A.a[t], o + k, fl);
        o[k++] = ' ';
        k += fmt_g(fict ? 0.0 : A.b[t], o + k, fl);
        o[k++] = ' ';
        if (KIND == R_VEL3) k += fmt_g(fict ? 0.0 : A.c[t], o + k, fl);
        else o[k++] = '0';
    } else if (KIND == R_F64) {
        k += fmt_g(A.a[t], o + k, fl);
    } else if (KIND == R_U8) {
        k += fmt_int((int)((const uint8_t*)A.ints)[t], o + k);
    } else {
        k += fmt_int(((const int*)A.ints)[t], o + k);
    }
    o[k++] = '\n';
    return k;
}

// pass 1: record lengths (the text itself is discarded; formatting twice is cheaper than a
// round trip of fixed-size slots through HBM)
template <int KIND>
__global__ void __launch_bounds__(256)
k_rec_len(RecArgs A, long long n, int* __restrict__ len, int* __restrict__ flag) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    char buf[kMaxRec];
    int fl = 0;
    len[t] = make_record<KIND>(A, t, buf, &fl);
    if (fl) atomicExch(flag, 1);
}

// pass 2: a block formats its 256 records into shared memory at their final relative positions
// (shifted so that shared index 0 maps to a 16-byte aligned global address) and writes the block's
// text with aligned 16-byte stores; only the ragged ends go out bytewise
template <int KIND>
__global__ void __launch_bounds__(256)
k_rec_write(RecArgs A, long long n, const long long* __restrict__ pos, char* __restrict__ text) {
    __shared__ __align__(16) char sbuf[256 * kMaxRec + 32];
    const long long t0 = (long long)blockIdx.x * 256, t = t0 + threadIdx.x;
    const long long t1 = t0 + 256 < n ? t0 + 256 : n;
    const long long base = pos[t0], end = pos[t1];
    const int mis = (int)((unsigned long long)(text + base) & 15ull);
    int fl = 0;
    if (t < n) make_record<KIND>(A, t, sbuf + mis + (int)(pos[t] - base), &fl);
    __syncthreads();
    const int total = (int)(end - base);
    char* g0 = text + base - mis;                  // 16-byte aligned, corresponds to sbuf[0]
    const int first_full = mis ? 16 : 0;           // first chunk that lies completely inside the text
    const int last = mis + total;                  // one past the last valid shared index
    const int full_end = last & ~15;
    for (int i = first_full + threadIdx.x * 16; i < full_end; i += 256 * 16)
        *(uint4*)(g0 + i) = *(const uint4*)(sbuf + i);
    if (mis) {
        const int head_end = last < 16 ? last : 16;
        for (int i = mis + threadIdx.x; i < head_end; i += 256) g0[i] = sbuf[i];
    }
    const int tail_begin = full_end > first_full ? full_end : (mis ? 16 : 0);
    for (int i = tail_begin + threadIdx.x; i < last; i += 256) g0[i] = sbuf[i];
}

// raw "%g" of an array (test hook): 16-byte zero-padded cells
__global__ void k_fmt_cells(const double* __restrict__ v, long long n, char* __restrict__ cells, int* __restrict__ flag) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    char* o = cells + t * 16;
    int fl = 0;
    int k = fmt_g(v[t], o, &fl);
    for (; k < 16; ++k) o[k] = 0;
    if (fl) atomicExch(flag, 1);
}

struct TextBuf {
    int* len = nullptr;
    long long* pos = nullptr;
    char* text = nullptr;
    int* flag = nullptr;
    long long n = 0;
    void release() {
        cudaFree(len); cudaFree(pos); cudaFree(text); cudaFree(flag);
        text = nullptr; len = flag = nullptr; pos = nullptr;
    }
    ~TextBuf() { release(); }   // every exit path of the callers frees the device buffers
};

// device scratch array that is freed on every exit path
template <typename T>
struct DevArray {
    T* p = nullptr;
    ~DevArray() { cudaFree(p); }
};

int upload_table(pdgpu_ctx* c) {
    if (c->device < 64 && g_table_uploaded[c->device]) return 0;
    CUDA_OK(cudaMemcpyToSymbol(d_pow10, kPow10, sizeof(kPow10)));
    if (c->device < 64) g_table_uploaded[c->device] = true;
    return 0;
}

int alloc_buf(pdgpu_ctx* c, TextBuf* b, long long n) {
    b->n = n;
    CUDA_OK(cudaMalloc(&b->len, sizeof(int) * n));
    CUDA_OK(cudaMalloc(&b->pos, sizeof(long long) * (n + 1)));
    CUDA_OK(cudaMalloc(&b->text, (size_t)n * kMaxRec + 64));
    CUDA_OK(cudaMalloc(&b->flag, sizeof(int)));
    CUDA_OK(cudaMemsetAsync(b->flag, 0, sizeof(int), c->stream));   // same stream as the kernels that raise it
    return 0;
}

enum ArrKind { A_VEL, A_F64, A_U8, A_I32 };
struct ArrSpec {
    const char* header;   // DataArray line of the reference
    ArrKind kind;
    const void* dev;      // device pointer (owned-node 0)
};

// formats one array into b->text; *bytes = length of the body
template <int KIND>
int format_kind(pdgpu_ctx* c, TextBuf* b, const RecArgs& A, long long* bytes) {
    const long long n = b->n;
    const unsigned g = nblocks(n, 256);
    LAUNCH(c, k_rec_len<KIND>, g, 256, 0, A, n, b->len, b->flag);
    long long total = 0;
    PD_TRY(pdscan::exclusive_scan(c, b->len, n, b->pos, &total));
    LAUNCH(c, k_rec_write<KIND>, g, 256, 0, A, n, b->pos, b->text);
    *bytes = total;
    return 0;
}

int format_array(pdgpu_ctx* c, TextBuf* b, const ArrSpec& a, long long* bytes) {
    RecArgs A;
    A.a = A.b = A.c = nullptr; A.type = nullptr; A.ints = nullptr;
    switch (a.kind) {
        case A_VEL: {
            const long long lo = c->own_lo;
            A.a = c->v[c->cur][0] + lo; A.b = c->v[c->cur][1] + lo; A.c = c->dim == 3 ? c->v[c->cur][2] + lo : nullptr;
            A.type = c->type + lo;
            return c->dim == 2 ? format_kind<R_VEL2>(c, b, A, bytes) : format_kind<R_VEL3>(c, b, A, bytes);
        }
        case A_F64: A.a = (const double*)a.dev; return format_kind<R_F64>(c, b, A, bytes);
        case A_U8: A.ints = a.dev; return format_kind<R_U8>(c, b, A, bytes);
        case A_I32: A.ints = a.dev; return format_kind<R_I32>(c, b, A, bytes);
    }
    return 0;
}

}  // namespace

extern "C" int pdgpu_format_g(pdgpu_ctx* c, const double* host_vals, long long n, char* host_cells16) {
    CHECK_CTX(c);
    if (!host_vals || !host_cells16 || n <= 0) PD_FAIL("pdgpu_format_g: bad arguments");
    PD_TRY(upload_table(c));
    double* dv = nullptr;
    char* dc = nullptr;
    int* df = nullptr;
    CUDA_OK(cudaMalloc(&dv, sizeof(double) * n));
    CUDA_OK(cudaMalloc(&dc, (size_t)16 * n));
    CUDA_OK(cudaMalloc(&df, sizeof(int)));
    CUDA_OK(cudaMemsetAsync(df, 0, sizeof(int), c->stream));
    CUDA_OK(cudaMemcpyAsync(dv, host_vals, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_fmt_cells, nblocks(n, 256), 256, 0, dv, n, dc, df);
    int flag = 0;
    CUDA_OK(cudaMemcpyAsync(host_cells16, dc, (size_t)16 * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(&flag, df, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    cudaFree(dv); cudaFree(dc); cudaFree(df);
    if (flag) PD_FAIL("pdgpu_format_g: a rounding decision could not be proven");
    return 0;
}

// VTKWriter::write (src/vtk_writer.cpp:16-146). grain_id / D_map are host-side arrays of the driver
// (never read by the solvers); nullptr writes -1 / 0 like an unset Fields.
extern "C" int pdgpu_vti_write(pdgpu_ctx* c, const char* path, const int* grain_id, const double* D_map,
                               long long* bytes_out, float* format_ms) {
    NEED_FIELDS(c);
    if (!path) PD_FAIL("pdgpu_vti_write: null path");
    if (c->nranks > 1) PD_FAIL("pdgpu_vti_write: slab contexts write per-rank pieces only (not implemented)");
    PD_TRY(pd_flush_wall_c(c));
    PD_TRY(upload_table(c));
    const long long n = c->own_hi - c->own_lo, lo = c->own_lo;
    TextBuf b;
    PD_TRY(alloc_buf(c, &b, n));
    DevArray<int> gid_buf;
    DevArray<double> dmap_buf, press_buf;
    CUDA_OK(cudaMalloc(&gid_buf.p, sizeof(int) * n));
    CUDA_OK(cudaMalloc(&dmap_buf.p, sizeof(double) * n));
    int* d_gid = gid_buf.p;
    double* d_dmap = dmap_buf.p;
    if (grain_id) CUDA_OK(cudaMemcpyAsync(d_gid, grain_id, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    else CUDA_OK(cudaMemsetAsync(d_gid, 0xFF, sizeof(int) * n, c->stream));
    if (D_map) CUDA_OK(cudaMemcpyAsync(d_dmap, D_map, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    else CUDA_OK(cudaMemsetAsync(d_dmap, 0, sizeof(double) * n, c->stream));
    // `pressure` is written as the reference computes it, B (pow(rho/rho_f, gamma) - 1) of the density the
    // last step started from (src/pd_ns.cpp:36-50,84), not from the Horner-evaluated shadow field the
    // tiled kernels keep (same value to ~1e-10 relative, which is visible in the 6th digit now and then)
    CUDA_OK(cudaMalloc(&press_buf.p, sizeof(double) * n));
    double* d_press = press_buf.p;
    PD_TRY(pd_enqueue_eos_to(c, c->p_input, lo, n, d_press));
    if (!c->stale_p_idx.empty()) {   // nodes dissolved since the last NS step keep their pre-dissolution pressure
        const PdConsts kc = pd_consts(c->cfg, c->dim);
        std::vector<double> pv(c->stale_p_idx.size());
        for (size_t t = 0; t < pv.size(); ++t) {
            double ratio = c->stale_p_rho[t] / c->cfg.rho_f;          // src/pd_ns.cpp:44-48
            ratio = std::min(std::max(ratio, 0.5), 2.0);
            pv[t] = kc.B_eos * (std::pow(ratio, c->cfg.gamma_eos) - 1.0);
        }
        for (size_t t = 0; t < pv.size(); ++t) {
            const long long l = c->stale_p_idx[t];
            if (l >= lo && l < lo + n)
                CUDA_OK(cudaMemcpyAsync(d_press + (l - lo), &pv[t], sizeof(double), cudaMemcpyHostToDevice, c->stream));
        }
        CUDA_OK(cudaStreamSynchronize(c->stream));   // pv lives on this stack frame
    }

    const int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) PD_FAIL("cannot open VTI file '%s'", path);
    const int nx = c->Nx, ny = c->Ny, nz = (c->dim == 3) ? c->Nz : 1;
    std::string head;
    {   // header, formatted by the same iostream rules as the reference (src/vtk_writer.cpp:40-53)
        std::ostringstream h;
        h << "<?xml version=\"1.0\"?>\n";
        h << "<VTKFile type=\"ImageData\" version=\"1.0\" byte_order=\"LittleEndian\">\n";
        h << "  <ImageData WholeExtent=\"0 " << nx - 1 << " 0 " << ny - 1 << " 0 " << nz - 1 << "\""
          << " Origin=\"" << c->origin[0] << " " << c->origin[1] << " " << ((c->dim == 3) ? c->origin[2] : 0.0) << "\""
          << " Spacing=\"" << c->cfg.dx << " " << c->cfg.dx << " " << c->cfg.dx << "\">\n";
        h << "    <Piece Extent=\"0 " << nx - 1 << " 0 " << ny - 1 << " 0 " << nz - 1 << "\">\n";
        h << "      <PointData Scalars=\"phase\" Vectors=\"velocity\">\n";
        head = h.str();
    }
    const ArrSpec arrs[] = {
        {"        <DataArray type=\"Float64\" Name=\"velocity\" NumberOfComponents=\"3\" format=\"ascii\">\n", A_VEL, nullptr},
        {"        <DataArray type=\"Float64\" Name=\"pressure\" format=\"ascii\">\n", A_F64, d_press},
        {"        <DataArray type=\"Float64\" Name=\"density\" format=\"ascii\">\n", A_F64, c->rho[c->cur] + lo},
        {"        <DataArray type=\"Float64\" Name=\"concentration\" format=\"ascii\">\n", A_F64, c->C[c->curC] + lo},
        {"        <DataArray type=\"UInt8\" Name=\"phase\" format=\"ascii\">\n", A_U8, c->phase + lo},
        {"        <DataArray type=\"UInt8\" Name=\"node_type\" format=\"ascii\">\n", A_U8, c->type + lo},
        {"        <DataArray type=\"Int32\" Name=\"grain_id\" format=\"ascii\">\n", A_I32, d_gid},
        {"        <DataArray type=\"Float64\" Name=\"D_map\" format=\"ascii\">\n", A_F64, d_dmap},
        {"        <DataArray type=\"UInt8\" Name=\"is_grain_boundary\" format=\"ascii\">\n", A_U8, c->is_gb + lo},
        {"        <DataArray type=\"UInt8\" Name=\"is_precipitate\" format=\"ascii\">\n", A_U8, c->is_precip + lo},
    };
    // The finished text leaves the device in chunks through a small pool of pinned buffers; writer
    // threads put every chunk at its final file offset (all offsets are known from the scans), so the
    // D2H copy of one chunk, the page-cache copies of others and the next formatting kernel overlap.
    pdio::IoPool* io = nullptr;
    int rc = pdio::io_pool(c, &io);
    long long total_bytes = 0;
    float ms_sum = 0.f;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    pdio::IoRun run(io, fd, c->device);
    off_t off = 0;
    auto put_small = [&](const std::string& t) {
        if (::pwrite(fd, t.data(), t.size(), off) != (ssize_t)t.size()) run.fail();
        off += (off_t)t.size();
    };
    if (!rc) put_small(head);
    for (const ArrSpec& a : arrs) {
        if (rc || run.failed()) break;
        long long bytes = 0;
        cudaEventRecord(e0, c->stream);
        rc = format_array(c, &b, a, &bytes);
        if (rc) break;
        cudaEventRecord(e1, c->stream);
        put_small(a.header);
        for (long long done = 0; done < bytes; done += (long long)pdio::kIoChunk) {
            const size_t len = (size_t)std::min<long long>((long long)pdio::kIoChunk, bytes - done);
            const int k = run.acquire();
            if (k < 0) break;
            if (cudaMemcpyAsync(io->buf[k], b.text + done, len, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                cudaEventRecord(io->ev[k], c->stream) != cudaSuccess) { run.fail(); break; }
            run.submit(k, len, off + (off_t)done);
        }
        off += (off_t)bytes;
        put_small("        </DataArray>\n");
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { run.fail(); break; }   // b.text is reused by the next array
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        ms_sum += ms;
        total_bytes += bytes;
    }
    int flag = 0;
    if (!rc) {
        cudaMemcpyAsync(&flag, b.flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        put_small("      </PointData>\n    </Piece>\n  </ImageData>\n</VTKFile>\n");
    }
    run.finish();
    ::close(fd);
    if (!rc && run.failed()) { rc = 1; pd_set_error("pdgpu_vti_write: writing '%s' failed", path); }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc) return rc;
    if (flag) PD_FAIL("pdgpu_vti_write: a rounding decision could not be proven");
    if (bytes_out) *bytes_out = total_bytes;
    if (format_ms) *format_ms = ms_sum;
    return 0;
}
