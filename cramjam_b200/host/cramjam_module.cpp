// cramjam_module.cpp — C++ (pybind11) host binding: the `cramjam` Python surface for the snappy / lz4 /
// zstd path, calling ONLY the extern "C" boundary of libcramjam_cuda.so (include/cramjam_cuda.h).
//
// It stands where the reference's Rust/pyo3 shim stands (no Rust toolchain exists in this image):
//   src/lib.rs:104-207   BytesType (Buffer | File | buffer-protocol object)       -> struct Input / deliver()
//   src/lib.rs:210-296   generic! (output allocation, in x out dispatch, GIL release) -> run_codec() / deliver()
//   src/io.rs:370-684    RustyBuffer "Buffer"                                       -> class Buffer
//   src/io.rs:40-172     RustyFile "File"                                           -> class File
//   src/exceptions.rs    CompressionError / DecompressionError                      -> py::register_exception
//   src/snappy.rs, src/lz4.rs, src/zstd.rs                                          -> submodules snappy, lz4, zstd
// Same function names, keyword arguments, return types and error behaviour; the codec work itself is
// done by the CUDA engine (no CPU codec path exists here).
#include <pybind11/pybind11.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cramjam_cuda.h"

namespace py = pybind11;

static PyObject* g_compression_error = nullptr;
static PyObject* g_decompression_error = nullptr;

[[noreturn]] static void raise(PyObject* type, const std::string& msg) {
    PyErr_SetString(type, msg.c_str());
    throw py::error_already_set();
}

// ---- engine context: one per process, created on first use (cj_ctx serialises its own calls) ----
static cj_ctx* engine() {
    static std::mutex mu;
    static cj_ctx* ctx = nullptr;
    std::lock_guard<std::mutex> g(mu);
    if (!ctx) {
        int dev = 0;
        if (const char* e = std::getenv("CRAMJAM_CUDA_DEVICE")) dev = std::atoi(e);
        if (cj_ctx_create(dev, &ctx) != CJ_OK) {
            ctx = nullptr;
            throw std::runtime_error(std::string("cramjam_b200: cannot create the CUDA engine: ") + cj_last_error());
        }
    }
    return ctx;
}

// Byte vector whose resize / sized construction leaves new bytes uninitialised: the engine overwrites the output
// right away, and value-initialising (plus first-touching) a large output on one thread cost ~100 ms per 256 MiB —
// the same cost the reference pays for `vec![0; len]` (src/lib.rs:217-220).  Where zeros are part of the contract
// (a fixed-capacity output that is returned whole) they are written explicitly.
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    using std::allocator<T>::allocator;
    template <class U> void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
using Bytes = std::vector<uint8_t, NoInitAlloc<uint8_t>>;

// ------------------------------------------------------------------------------------------------
// Buffer (src/io.rs:370-684)
// ------------------------------------------------------------------------------------------------
class Buffer {
public:
    Bytes own;
    py::object view_ref;  // non-null => this Buffer is a view over another object's memory
    uint8_t* vptr = nullptr;
    size_t vlen = 0;
    size_t pos = 0;

    // Pinned staging (north_star: "Buffer gains a pinned-host/device staging path", reference src/io.rs:370-375): with
    // pinned=True the owned storage is kept page-locked (cj_host_register over the vector's capacity), so that codec calls
    // between two pinned Buffers run as CJ_PINNED — the engine DMAs straight from / to these pages, no staging copy.
    bool pinned = false;
    uint8_t* reg_ptr = nullptr;   // start of the range that is registered right now

    Buffer() = default;
    explicit Buffer(Bytes&& v) : own(std::move(v)) {}
    Buffer(const Buffer&) = delete;
    ~Buffer() { unregister(); }

    void unregister() {
        if (reg_ptr) {
            cj_host_unregister(engine(), reg_ptr);
            reg_ptr = nullptr;
        }
    }
    // call before anything that may reallocate the vector: registered pages must not be freed underneath the driver
    void reserve(size_t n) {
        if (n > own.capacity()) {
            unregister();
            own.reserve(std::max(n, own.capacity() + own.capacity() / 2));
        }
    }
    // true when the storage is page-locked and may be handed to the engine as CJ_PINNED
    bool dma_ready() {
        if (!pinned || is_view() || own.capacity() == 0) return false;
        if (reg_ptr != own.data()) {
            unregister();
            if (cj_host_register(engine(), own.data(), own.capacity()) == CJ_OK) reg_ptr = own.data();
        }
        return reg_ptr == own.data();
    }

    bool is_view() const { return !view_ref.is_none() && view_ref.ptr() != nullptr; }
    uint8_t* data() { return is_view() ? vptr : own.data(); }
    size_t size() const { return is_view() ? vlen : own.size(); }

    // src/io.rs:452-485: the object a view refers to may have been resized or moved
    void realign() {
        if (!is_view()) return;
        Py_buffer b;
        if (PyObject_GetBuffer(view_ref.ptr(), &b, PyBUF_CONTIG_RO) != 0) throw py::error_already_set();
        vptr = (uint8_t*)b.buf;
        vlen = (size_t)b.len;
        PyBuffer_Release(&b);
        if (pos > vlen) pos = vlen;
    }
};

// ------------------------------------------------------------------------------------------------
// File (src/io.rs:40-172)
// ------------------------------------------------------------------------------------------------
class File {
public:
    std::string path;
    FILE* f = nullptr;
    bool append = false;
    File(const std::string& p, py::object read, py::object write, py::object truncate, py::object append_) : path(p) {
        const bool r = read.is_none() ? true : read.cast<bool>();
        const bool w = write.is_none() ? true : write.cast<bool>();
        const bool t = truncate.is_none() ? false : truncate.cast<bool>();
        append = append_.is_none() ? false : append_.cast<bool>();
        // create if missing, open if present (OpenOptions::create(true))
        FILE* probe = std::fopen(p.c_str(), "ab");
        if (!probe) { PyErr_SetFromErrnoWithFilename(PyExc_OSError, p.c_str()); throw py::error_already_set(); }
        std::fclose(probe);
        const char* mode = append ? (r ? "a+b" : "ab") : (w ? (t ? "w+b" : "r+b") : "rb");
        f = std::fopen(p.c_str(), mode);
        if (!f) { PyErr_SetFromErrnoWithFilename(PyExc_OSError, p.c_str()); throw py::error_already_set(); }
    }
    ~File() { if (f) std::fclose(f); }
    File(const File&) = delete;
    size_t len() {
        std::fflush(f);
        const long cur = std::ftell(f);
        std::fseek(f, 0, SEEK_END);
        const long e = std::ftell(f);
        std::fseek(f, cur, SEEK_SET);
        return (size_t)e;
    }
    size_t tell() { return (size_t)std::ftell(f); }
    Bytes read_to_end() {
        Bytes out;
        const size_t total = len(), cur = tell();
        if (total > cur) {
            out.resize(total - cur);
            const size_t got = std::fread(out.data(), 1, out.size(), f);
            out.resize(got);
        }
        return out;
    }
    void write_all(const uint8_t* p, size_t n) {
        if (n && std::fwrite(p, 1, n, f) != n) { PyErr_SetFromErrno(PyExc_OSError); throw py::error_already_set(); }
        std::fflush(f);
    }
};

// ------------------------------------------------------------------------------------------------
// BytesType (src/lib.rs:104-207): a borrowed view of any accepted input / output object
// ------------------------------------------------------------------------------------------------
struct PyBuf {  // src/io.rs:177-335 PythonBuffer: PyObject_GetBuffer(PyBUF_CONTIG_RO), released on drop
    Py_buffer b{};
    bool held = false;
    explicit PyBuf(py::handle h) {
        if (PyObject_GetBuffer(h.ptr(), &b, PyBUF_CONTIG_RO) != 0) throw py::error_already_set();
        held = true;
        if (!PyBuffer_IsContiguous(&b, 'C')) {
            PyBuffer_Release(&b);
            held = false;
            raise(PyExc_BufferError, "Buffer is not C contiguous");
        }
    }
    ~PyBuf() { if (held) PyBuffer_Release(&b); }
    PyBuf(const PyBuf&) = delete;
};

// A device-resident array (anything exposing __cuda_array_interface__: torch / cupy / numba CUDA arrays).  Returns false
// when `h` is not one; raises BufferError when it is but is not contiguous.
static bool cuda_array(py::handle h, uint8_t** ptr, size_t* nbytes, bool* readonly) {
    if (py::isinstance<Buffer>(h) || PyObject_CheckBuffer(h.ptr())) return false;
    if (!PyObject_HasAttrString(h.ptr(), "__cuda_array_interface__")) return false;
    py::dict d = h.attr("__cuda_array_interface__").cast<py::dict>();
    if (d.contains("strides") && !d["strides"].is_none()) raise(PyExc_BufferError, "device array is not C contiguous");
    py::tuple data = d["data"].cast<py::tuple>();
    *ptr = reinterpret_cast<uint8_t*>(data[0].cast<uintptr_t>());
    *readonly = data[1].cast<bool>();
    const std::string ts = d["typestr"].cast<std::string>();
    size_t n = (size_t)std::atoi(ts.c_str() + 2);
    for (py::handle e : d["shape"].cast<py::tuple>()) n *= e.cast<size_t>();
    *nbytes = n;
    return true;
}

struct Input {
    const uint8_t* p = nullptr;
    size_t n = 0;
    bool dma = false;   // host pages that are page-locked (a pinned Buffer)
    std::unique_ptr<PyBuf> pb;
    Bytes tmp;  // File contents (read from the current position to the end)
    explicit Input(py::handle h) {
        if (py::isinstance<Buffer>(h)) {
            Buffer& b = h.cast<Buffer&>();
            b.realign();
            p = b.data();  // whole buffer, cursor ignored (src/lib.rs:229-233, src/io.rs:389-392)
            n = b.size();
            dma = b.dma_ready();
        } else if (py::isinstance<File>(h)) {
            tmp = h.cast<File&>().read_to_end();
            p = tmp.data();
            n = tmp.size();
        } else {
            pb.reset(new PyBuf(h));
            p = (const uint8_t*)pb->b.buf;
            n = (size_t)pb->b.len;
        }
    }
};

// Writes `n` produced bytes into an output BytesType the way generic! does (src/lib.rs:239-295):
// File -> at its position; Buffer -> at its cursor (grows; a view cannot); buffer-protocol object ->
// from its start, fixed capacity, overflow is an error.
static size_t deliver(py::handle out, const uint8_t* p, size_t n, PyObject* err_type) {
    if (py::isinstance<Buffer>(out)) {
        Buffer& b = out.cast<Buffer&>();
        b.realign();
        if (b.is_view()) {
            if (n > b.vlen - b.pos) raise(err_type, "failed to write whole buffer");
            std::memcpy(b.vptr + b.pos, p, n);
        } else {
            const size_t old = b.own.size();
            if (b.pos + n > old) { b.reserve(b.pos + n); b.own.resize(b.pos + n); }
            if (b.pos > old) std::memset(b.own.data() + old, 0, b.pos - old);  // a write past the end zero-fills the gap (Cursor<Vec<u8>>)
            if (n) std::memcpy(b.own.data() + b.pos, p, n);
        }
        b.pos += n;
        return n;
    }
    if (py::isinstance<File>(out)) {
        out.cast<File&>().write_all(p, n);
        return n;
    }
    PyBuf pb(out);
    if (n > (size_t)pb.b.len) raise(err_type, "failed to write whole buffer");
    if (n) std::memcpy(pb.b.buf, p, n);  // writes through read-only exports too, like the reference on CPython (src/io.rs:210-234)
    return n;
}

static size_t output_capacity(py::handle out) {  // only meaningful for fixed-capacity outputs
    if (py::isinstance<Buffer>(out) || py::isinstance<File>(out)) return SIZE_MAX;
    PyBuf pb(out);
    return (size_t)pb.b.len;
}

// ---- codec calls (GIL released, like py.allow_threads in generic!) ----
static Bytes do_compress(cj_codec codec, const uint8_t* p, size_t n, int level = -1, int accel = 1) {
    Bytes out(cj_compress_bound(codec, n));
    size_t written = 0;
    cj_params prm{level, accel, 0};
    int rc;
    std::string msg;
    {
        py::gil_scoped_release rel;
        rc = cj_compress(engine(), codec, p, n, out.data(), out.size(), &written, &prm);
        if (rc) msg = cj_last_error();
    }
    if (rc) raise(g_compression_error, msg);
    out.resize(written);
    return out;
}

// cap == SIZE_MAX: size the output from the stream's own headers and shrink to what was produced;
// otherwise exactly `cap` bytes are allocated and kept (the reference does not truncate, benchmarks/README.md:24-28).
static Bytes do_decompress(cj_codec codec, const uint8_t* p, size_t n, size_t cap, size_t* produced = nullptr) {
    bool shrink = false;
    int rc;
    std::string msg;
    if (cap == SIZE_MAX) {
        size_t b = 0;
        rc = cj_decompress_bound(codec, p, n, &b);
        if (rc) raise(g_decompression_error, cj_last_error());
        cap = b;
        shrink = true;
    }
    Bytes out(cap);
    size_t written = 0;
    {
        py::gil_scoped_release rel;
        rc = cj_decompress(engine(), codec, p, n, out.data(), out.size(), &written);
        if (rc) msg = cj_last_error();
    }
    if (rc) raise(g_decompression_error, msg);
    if (shrink) out.resize(written);
    else if (written < out.size()) std::memset(out.data() + written, 0, out.size() - written);  // `vec![0; n]` returned whole
    if (produced) *produced = written;
    return out;
}

static size_t opt_size(const py::object& o) { return o.is_none() ? SIZE_MAX : o.cast<size_t>(); }

static py::object make_buffer(Bytes v) { return py::cast(new Buffer(std::move(v)), py::return_value_policy::take_ownership); }

// ---- the generic compress / decompress / *_into quartet shared by the three variants ----
// `output_len` is a size hint in the reference, not a capacity: generic! builds `vec![0; len]` and writes through a
// Cursor over that Vec, which grows (src/lib.rs:217-233).  The result is therefore max(len, produced) bytes long, zero
// padded behind what was produced, and an undersized hint still succeeds.
static void pad_to_hint(Bytes& v, size_t hint) {
    if (hint != SIZE_MAX && hint > v.size()) {
        const size_t old = v.size();
        v.resize(hint);
        std::memset(v.data() + old, 0, hint - old);
    }
}
static py::object generic_compress(cj_codec codec, py::handle data, int level, const py::object& output_len = py::none()) {
    Input in(data);
    Bytes out = do_compress(codec, in.p, in.n, level);
    pad_to_hint(out, opt_size(output_len));
    return make_buffer(std::move(out));
}
static py::object generic_decompress(cj_codec codec, py::handle data, const py::object& output_len) {
    Input in(data);
    Bytes out = do_decompress(codec, in.p, in.n, SIZE_MAX);
    pad_to_hint(out, opt_size(output_len));
    return make_buffer(std::move(out));
}
// Both sides device-resident: the call runs as CJ_DEVICE, nothing crosses PCIe.  Mixed pairs are refused (the caller decides
// where a copy happens).  Returns false when neither side is a device array.
static bool device_into(cj_codec codec, bool compress, py::handle input, py::handle output, int level, int accel, size_t* written) {
    uint8_t *sp = nullptr, *dp = nullptr;
    size_t sn = 0, dn = 0;
    bool sro = false, dro = false;
    const bool sd = cuda_array(input, &sp, &sn, &sro), dd = cuda_array(output, &dp, &dn, &dro);
    if (!sd && !dd) return false;
    PyObject* err = compress ? g_compression_error : g_decompression_error;
    if (sd != dd) raise(PyExc_ValueError, "device arrays and host buffers cannot be mixed in one call: pass both sides as device arrays (or copy one side first)");
    if (dro) raise(PyExc_BufferError, "output device array is read-only");
    cj_params prm{level, accel, 0};
    int rc;
    std::string msg;
    {
        py::gil_scoped_release rel;
        rc = compress ? cj_compress_ex(engine(), codec, CJ_DEVICE, sp, sn, dp, dn, written, &prm) : cj_decompress_ex(engine(), codec, CJ_DEVICE, sp, sn, dp, dn, written);
        if (rc) msg = cj_last_error();
    }
    if (rc) raise(err, msg);
    return true;
}

// Output is an owned Buffer: the engine writes straight into its storage at the cursor (no intermediate vector), as
// CJ_PINNED when both sides are page-locked.  `cap` = the most the call may produce.
static size_t into_buffer(cj_codec codec, bool compress, const Input& in, Buffer& b, size_t cap, int level, int accel) {
    const size_t old = b.own.size();
    b.reserve(b.pos + cap);
    if (b.pos + cap > old) b.own.resize(b.pos + cap);
    if (b.pos > old) std::memset(b.own.data() + old, 0, b.pos - old);
    const cj_mem where = (in.dma && b.dma_ready()) ? CJ_PINNED : CJ_HOST;
    cj_params prm{level, accel, 0};
    size_t written = 0;
    int rc;
    std::string msg;
    {
        py::gil_scoped_release rel;
        rc = compress ? cj_compress_ex(engine(), codec, where, in.p, in.n, b.own.data() + b.pos, cap, &written, &prm)
                      : cj_decompress_ex(engine(), codec, where, in.p, in.n, b.own.data() + b.pos, cap, &written);
        if (rc) msg = cj_last_error();
    }
    if (rc) {
        b.own.resize(old);
        raise(compress ? g_compression_error : g_decompression_error, msg);
    }
    b.own.resize(std::max(old, b.pos + written));
    b.pos += written;
    return written;
}

static size_t generic_compress_into(cj_codec codec, py::handle input, py::handle output, int level) {
    size_t dw = 0;
    if (device_into(codec, true, input, output, level, 1, &dw)) return dw;
    Input in(input);
    if (py::isinstance<Buffer>(output) && !output.cast<Buffer&>().is_view())
        return into_buffer(codec, true, in, output.cast<Buffer&>(), cj_compress_bound(codec, in.n), level, 1);
    Bytes c = do_compress(codec, in.p, in.n, level);
    return deliver(output, c.data(), c.size(), g_compression_error);
}
static size_t generic_decompress_into(cj_codec codec, py::handle input, py::handle output) {
    size_t dw = 0;
    if (device_into(codec, false, input, output, -1, 1, &dw)) return dw;
    Input in(input);
    if (py::isinstance<Buffer>(output) && !output.cast<Buffer&>().is_view()) {
        size_t bound = 0;
        if (cj_decompress_bound(codec, in.p, in.n, &bound)) raise(g_decompression_error, cj_last_error());
        return into_buffer(codec, false, in, output.cast<Buffer&>(), bound, -1, 1);
    }
    const size_t cap = output_capacity(output);
    size_t produced = 0;
    Bytes d;
    if (cap == SIZE_MAX) {
        d = do_decompress(codec, in.p, in.n, SIZE_MAX, &produced);
    } else {
        // fixed-capacity output: decode against exactly that capacity; too small is an error, never a truncation
        d = do_decompress(codec, in.p, in.n, cap, &produced);
    }
    return deliver(output, d.data(), produced, g_decompression_error);
}

// ---- batch entry points (not in the reference's Python surface; SURVEY.md 7 step 7 / north_star: "many independent buffers"):
//      a list of independent buffers goes to the engine as ONE cj_*_batch call — thousands of units per launch instead of
//      one launch per buffer — and comes back as a list of Buffers.  `caps[i]` = capacity of output i.
static py::list run_batch_call(cj_codec codec, bool compress, std::vector<std::unique_ptr<Input>>& ins, const std::vector<size_t>& skip,
                               const std::vector<size_t>& caps, int level, int accel) {
    const size_t n = ins.size();
    std::vector<Bytes> outs(n);
    std::vector<uint64_t> so(n), sl(n), dof(n), dc(n), dl(n, 0);
    std::vector<int32_t> st(n, 0);
    static uint8_t dummy[16];
    uintptr_t sbase = UINTPTR_MAX, dbase = UINTPTR_MAX;
    for (size_t i = 0; i < n; i++) {
        outs[i] = Bytes(std::max<size_t>(caps[i], 1));
        const uintptr_t sp = (uintptr_t)(ins[i]->n > skip[i] ? ins[i]->p + skip[i] : dummy);
        sbase = std::min(sbase, sp);
        dbase = std::min(dbase, (uintptr_t)outs[i].data());
    }
    for (size_t i = 0; i < n; i++) {
        const uintptr_t sp = (uintptr_t)(ins[i]->n > skip[i] ? ins[i]->p + skip[i] : dummy);
        so[i] = sp - sbase;
        sl[i] = ins[i]->n > skip[i] ? ins[i]->n - skip[i] : 0;
        dof[i] = (uintptr_t)outs[i].data() - dbase;
        dc[i] = caps[i];
    }
    cj_batch b;
    b.n = n;
    b.src_base = (const void*)sbase; b.src_off = so.data(); b.src_len = sl.data();
    b.dst_base = (void*)dbase; b.dst_off = dof.data(); b.dst_cap = dc.data();
    b.dst_len = dl.data(); b.status = st.data();
    cj_params prm{level, accel, 0};
    int rc = 0;
    std::string msg;
    if (n) {
        py::gil_scoped_release rel;
        rc = compress ? cj_compress_batch(engine(), codec, CJ_HOST, &b, &prm) : cj_decompress_batch(engine(), codec, CJ_HOST, &b);
        if (rc) msg = cj_last_error();
    }
    PyObject* err = compress ? g_compression_error : g_decompression_error;
    if (rc) raise(err, msg);
    py::list result;
    for (size_t i = 0; i < n; i++) {
        if (st[i] != CJ_OK) raise(err, "unit " + std::to_string(i) + ": " + cj_status_string(st[i]));
        outs[i].resize((size_t)dl[i]);
        result.append(make_buffer(std::move(outs[i])));
    }
    return result;
}

static py::list generic_batch(cj_codec codec, bool compress, py::iterable inputs, int level) {
    std::vector<std::unique_ptr<Input>> ins;
    for (py::handle h : inputs) ins.emplace_back(new Input(h));
    std::vector<size_t> skip(ins.size(), 0), caps(ins.size());
    for (size_t i = 0; i < ins.size(); i++) {
        if (compress) caps[i] = cj_compress_bound(codec, ins[i]->n);
        else if (cj_decompress_bound(codec, ins[i]->p, ins[i]->n, &caps[i])) raise(g_decompression_error, "unit " + std::to_string(i) + ": " + cj_last_error());
    }
    return run_batch_call(codec, compress, ins, skip, caps, level, 1);
}

// ---- streaming classes (src/io.rs:760-814, src/lib.rs:299-395): host-side accumulation feeding the engine ----
class Compressor {
public:
    cj_codec codec;
    int level;
    Bytes pending;
    bool finished = false;
    Compressor(cj_codec c, int lvl) : codec(c), level(lvl) {}
    size_t compress(py::handle input) {
        if (finished) raise(g_compression_error, "Compressor looks to have been consumed via `finish()`. please create a new compressor instance.");
        Input in(input);
        pending.insert(pending.end(), in.p, in.p + in.n);
        return in.n;
    }
    bool emitted_any = false;
    py::object emit() {  // a complete stream / frame per call; concatenated streams are legal in all three formats
        Bytes out = do_compress(codec, pending.data(), pending.size(), level);
        pending.clear();
        emitted_any = true;
        return make_buffer(std::move(out));
    }
    py::object flush() {
        if (finished || pending.empty()) return make_buffer({});
        return emit();
    }
    py::object finish() {
        if (finished) return make_buffer({});
        finished = true;
        if (pending.empty() && emitted_any) return make_buffer({});
        return emit();  // also the "nothing was ever written" case: a valid empty stream
    }
};

class Decompressor {
public:
    cj_codec codec;
    Bytes acc;
    bool finished = false;
    explicit Decompressor(cj_codec c) : codec(c) {}
    void check() const { if (finished) raise(g_decompression_error, "Appears `finish()` was called on this instance"); }
    size_t decompress(py::handle input) {
        check();
        Input in(input);
        Bytes d = do_decompress(codec, in.p, in.n, SIZE_MAX);
        acc.insert(acc.end(), d.begin(), d.end());
        return d.size();
    }
    py::object flush() {
        check();
        Bytes out;
        out.swap(acc);
        return make_buffer(std::move(out));
    }
    py::object finish() {
        check();
        finished = true;
        Bytes out;
        out.swap(acc);
        return make_buffer(std::move(out));
    }
    size_t len() const { return finished ? 0 : acc.size(); }
};

// pybind11 registers one Python class per C++ type, so every variant gets its own tagged subtype.
template <int Tag> struct CompressorT : Compressor { using Compressor::Compressor; };
template <int Tag> struct DecompressorT : Decompressor { using Decompressor::Decompressor; };

// kind: 0 = no constructor arguments (snappy), 1 = level (zstd), 2 = level, content_checksum, block_linked (lz4)
template <int Tag, class M>
static void add_stream_classes(M& m, cj_codec codec, int kind, int default_level) {
    using C = CompressorT<Tag>;
    using D = DecompressorT<Tag>;
    auto comp = py::class_<C>(m, "Compressor");
    if (kind == 1)
        comp.def(py::init([codec, default_level](py::object level) { return new C(codec, level.is_none() ? default_level : level.cast<int>()); }), py::arg("level") = py::none());
    else if (kind == 2)
        comp.def(py::init([codec, default_level](py::object level, py::object, py::object) { return new C(codec, level.is_none() ? default_level : level.cast<int>()); }),
                 py::arg("level") = py::none(), py::arg("content_checksum") = py::none(), py::arg("block_linked") = py::none());
    else
        comp.def(py::init([codec]() { return new C(codec, -1); }));
    comp.def("compress", [](C& c, py::handle input) { return c.compress(input); }, py::arg("input"))
        .def("flush", [](C& c) { return c.flush(); })
        .def("finish", [](C& c) { return c.finish(); });
    py::class_<D>(m, "Decompressor")
        .def(py::init([codec]() { return new D(codec); }))
        .def("decompress", [](D& d, py::handle input) { return d.decompress(input); }, py::arg("input"))
        .def("flush", [](D& d) { return d.flush(); })
        .def("finish", [](D& d) { return d.finish(); })
        .def("len", [](D& d) { return d.len(); })
        .def("__len__", [](D& d) { return d.len(); })
        .def("__bool__", [](D& d) { return !d.finished && !d.acc.empty(); })
        .def("__repr__", [](D& d) { return "Decompressor<len=" + std::to_string(d.len()) + ">"; });
}

// ---- lz4 block helpers (src/lz4.rs:78-229) ----
// `compression=Some(n)` is CompressionMode::HIGHCOMPRESSION(n) in the reference (src/lz4.rs:113-131): passed as the level
static Bytes lz4_block_compress(const uint8_t* p, size_t n, bool store_size, int accel, int compression = -1) {
    Bytes c = do_compress(CJ_LZ4_BLOCK, p, n, compression >= 0 ? std::max(compression, 3) : -1, accel);
    if (!store_size) return c;
    Bytes out(c.size() + 4);
    const uint32_t sz = (uint32_t)n;  // 4-byte little-endian uncompressed-size prefix (lz4::block, prepend_size)
    std::memcpy(out.data(), &sz, 4);
    if (!c.empty()) std::memcpy(out.data() + 4, c.data(), c.size());
    return out;
}

// lz4::block::decompress_to_buffer semantics: size_prepended reads the prefix, else the capacity is the size.
static bool lz4_block_decompress_into(const uint8_t* p, size_t n, uint8_t* out, size_t out_len, bool size_prepended, size_t* written, std::string* err) {
    size_t size = out_len;
    if (size_prepended) {
        if (n < 4) { *err = "Source buffer must at least contain size prefix."; return false; }
        int32_t s;
        std::memcpy(&s, p, 4);
        if (s < 0) { *err = "Parsed size prefix in buffer must not be negative."; return false; }
        size = (size_t)s;
        p += 4;
        n -= 4;
    }
    if (size > 0x7E000000u) { *err = "Given size parameter is too big"; return false; }
    if (size > out_len) { *err = "buffer isn't large enough to hold decompressed data"; return false; }
    int rc;
    {
        py::gil_scoped_release rel;
        rc = cj_decompress(engine(), CJ_LZ4_BLOCK, p, n, out, size, written);
        if (rc) *err = std::string("Decompression failed. Input invalid or too long? (") + cj_last_error() + ")";
    }
    return rc == 0;
}

PYBIND11_MODULE(cramjam, m) {
    m.doc() = "cramjam-compatible snappy / lz4 / zstd surface backed by the B200 CUDA engine (libcramjam_cuda.so)";
    m.attr("__version__") = "2.12.0+b200.0";
    g_compression_error = PyErr_NewException("cramjam.CompressionError", PyExc_Exception, nullptr);
    g_decompression_error = PyErr_NewException("cramjam.DecompressionError", PyExc_Exception, nullptr);
    m.attr("CompressionError") = py::reinterpret_borrow<py::object>(g_compression_error);
    m.attr("DecompressionError") = py::reinterpret_borrow<py::object>(g_decompression_error);

    // ------------------------------------------------------------------ Buffer
    py::class_<Buffer>(m, "Buffer", py::buffer_protocol())
        .def(py::init([](py::object data, py::object copy, bool pinned) {
                 auto* b = new Buffer();
                 b->pinned = pinned;
                 if (data.is_none()) return b;
                 const bool do_copy = copy.is_none() ? true : copy.cast<bool>();
                 try {
                     if (do_copy) {
                         if (py::isinstance<Buffer>(data)) {
                             // reads from the source's cursor to its end, advancing it (bytestype.read_to_end)
                             Buffer& s = data.cast<Buffer&>();
                             s.realign();
                             b->own.assign(s.data() + s.pos, s.data() + s.size());
                             s.pos = s.size();
                         } else {
                             Input in(data);
                             b->own.assign(in.p, in.p + in.n);
                         }
                     } else {
                         PyBuf pb(data);
                         b->view_ref = data;
                         b->vptr = (uint8_t*)pb.b.buf;
                         b->vlen = (size_t)pb.b.len;
                     }
                 } catch (...) {
                     delete b;
                     throw;
                 }
                 return b;
             }),
             py::arg("data") = py::none(), py::arg("copy") = py::none(), py::arg("pinned") = false)
        .def_property_readonly("pinned", [](Buffer& b) { return b.pinned; },
                               "True when the Buffer keeps its storage page-locked for direct DMA (not in the reference: the B200 staging path)")
        .def("reserve", [](Buffer& b, size_t n) {
            if (b.is_view()) raise(PyExc_OSError, "Cannot reserve on unowned buffer");
            b.reserve(n);
            return b.own.capacity();
        }, py::arg("n"), "Grow the capacity to at least n bytes (a pinned Buffer registers its pages once for that capacity)")
        .def_buffer([](Buffer& b) {
            b.realign();
            return py::buffer_info(b.data(), 1, "B", 1, {(py::ssize_t)b.size()}, {(py::ssize_t)1}, /*readonly=*/true);
        })
        .def("get_view_reference", [](Buffer& b) -> py::object { return b.is_view() ? b.view_ref : py::none(); })
        .def("get_view_reference_count", [](Buffer& b) -> py::object {
            if (!b.is_view()) return py::none();
            return py::int_((py::ssize_t)Py_REFCNT(b.view_ref.ptr()));
        })
        .def("len", [](Buffer& b) { b.realign(); return b.size(); })
        .def("write", [](Buffer& b, py::handle input) {
            b.realign();
            Input in(input);
            if (b.is_view() && in.n > b.vlen - b.pos) raise(PyExc_OSError, "Too much to write on view");
            return deliver(py::cast(&b, py::return_value_policy::reference), in.p, in.n, PyExc_OSError);
        }, py::arg("input"))
        .def("read", [](Buffer& b, py::object n_bytes) {
            b.realign();
            size_t remaining = b.size() - b.pos, n = remaining;
            if (!n_bytes.is_none()) {
                const py::ssize_t v = n_bytes.cast<py::ssize_t>();
                if (v >= 0) n = std::min<size_t>((size_t)v, remaining);
            }
            py::bytes out((const char*)b.data() + b.pos, n);
            b.pos += n;
            return out;
        }, py::arg("n_bytes") = py::none())
        .def("readinto", [](Buffer& b, py::handle output) {
            b.realign();
            size_t n = b.size() - b.pos;
            const size_t cap = output_capacity(output);
            if (cap != SIZE_MAX) n = std::min(n, cap);
            const size_t w = deliver(output, b.data() + b.pos, n, PyExc_OSError);
            b.pos += w;
            return w;
        }, py::arg("output"))
        .def("seek", [](Buffer& b, py::ssize_t position, py::object whence_o) {
            b.realign();
            const size_t whence = whence_o.is_none() ? 0 : whence_o.cast<size_t>();
            const py::ssize_t len = (py::ssize_t)b.size(), cur = (py::ssize_t)b.pos;
            py::ssize_t target;
            if (whence == 0) target = position;
            else if (whence == 1) target = cur + position;
            else if (whence == 2) target = len + position;
            else raise(PyExc_ValueError, "whence should be one of 0: seek from start, 1: seek from current, or 2: seek from end");
            if (b.is_view() && (target > len || target < 0))
                raise(PyExc_OSError, "Bad seek: cannot seek outside bounds of unowned buffer, which has length of " + std::to_string(len) + ".");
            if (target < 0) raise(PyExc_OSError, "invalid seek to a negative or overflowing position");
            b.pos = (size_t)target;
            return b.pos;
        }, py::arg("position"), py::arg("whence") = py::none())
        .def("seekable", [](Buffer&) { return true; })
        .def("tell", [](Buffer& b) { b.realign(); return b.pos; })
        .def("set_len", [](Buffer& b, size_t size) {
            if (b.is_view()) raise(PyExc_OSError, "Cannot set length on unowned buffer");
            b.reserve(size);
            b.own.resize(size, 0);
        }, py::arg("size"))
        .def("truncate", [](Buffer& b) {
            if (b.is_view()) raise(PyExc_OSError, "Cannot truncate unowned buffer");
            b.own.clear();
            b.pos = 0;
        })
        .def("__len__", [](Buffer& b) { b.realign(); return b.size(); })
        .def("__bool__", [](Buffer& b) { b.realign(); return b.size() > 0; })
        .def("__repr__", [](Buffer& b) { b.realign(); return "cramjam.Buffer<len=" + std::to_string(b.size()) + ">"; })
        .def("__contains__", [](Buffer& b, py::handle x) {
            Input in(x);
            b.realign();
            if (in.n == 0 || in.n > b.size()) return false;  // slice::windows(0) panics in the reference; empty needle is simply absent here
            for (size_t i = 0; i + in.n <= b.size(); i++)
                if (std::memcmp(b.data() + i, in.p, in.n) == 0) return true;
            return false;
        })
        .def("__eq__", [](Buffer& a, py::handle other) {
            if (!py::isinstance<Buffer>(other)) return false;
            Buffer& b = other.cast<Buffer&>();
            a.realign();
            b.realign();
            return a.size() == b.size() && a.pos == b.pos && (a.size() == 0 || std::memcmp(a.data(), b.data(), a.size()) == 0);
        })
        .def("__ne__", [](Buffer& a, py::handle other) {
            if (!py::isinstance<Buffer>(other)) return true;
            Buffer& b = other.cast<Buffer&>();
            a.realign();
            b.realign();
            return !(a.size() == b.size() && a.pos == b.pos && (a.size() == 0 || std::memcmp(a.data(), b.data(), a.size()) == 0));
        });

    // ------------------------------------------------------------------ File
    py::class_<File>(m, "File")
        .def(py::init<const std::string&, py::object, py::object, py::object, py::object>(), py::arg("path"), py::arg("read") = py::none(),
             py::arg("write") = py::none(), py::arg("truncate") = py::none(), py::arg("append") = py::none())
        .def("write", [](File& f, py::handle input) {
            if (py::isinstance<File>(input)) {
                Bytes d = input.cast<File&>().read_to_end();
                f.write_all(d.data(), d.size());
                return d.size();
            }
            if (py::isinstance<Buffer>(input)) {  // copy(&mut buf.inner, output): from the source cursor to its end
                Buffer& s = input.cast<Buffer&>();
                s.realign();
                const size_t n = s.size() - s.pos;
                f.write_all(s.data() + s.pos, n);
                s.pos = s.size();
                return n;
            }
            Input in(input);
            f.write_all(in.p, in.n);
            return in.n;
        }, py::arg("input"))
        .def("read", [](File& f, py::object n_bytes) {
            Bytes out;
            if (n_bytes.is_none()) out = f.read_to_end();
            else {
                out.assign(n_bytes.cast<size_t>(), 0);  // PyBytes::new_with(n): short reads leave zero padding, as in the reference
                const size_t got = std::fread(out.data(), 1, out.size(), f.f);
                (void)got;
            }
            return py::bytes((const char*)out.data(), out.size());
        }, py::arg("n_bytes") = py::none())
        .def("readinto", [](File& f, py::handle output) {
            const size_t cap = output_capacity(output);
            Bytes d;
            if (cap == SIZE_MAX) d = f.read_to_end();
            else {
                d.resize(cap);
                d.resize(std::fread(d.data(), 1, cap, f.f));
            }
            return deliver(output, d.data(), d.size(), PyExc_OSError);
        }, py::arg("output"))
        .def("seek", [](File& f, py::ssize_t position, py::object whence_o) {
            const size_t whence = whence_o.is_none() ? 0 : whence_o.cast<size_t>();
            if (whence > 2) raise(PyExc_ValueError, "whence should be one of 0: seek from start, 1: seek from current, or 2: seek from end");
            std::fflush(f.f);
            if (std::fseek(f.f, (long)position, whence == 0 ? SEEK_SET : (whence == 1 ? SEEK_CUR : SEEK_END)) != 0) {
                PyErr_SetFromErrno(PyExc_OSError);
                throw py::error_already_set();
            }
            return f.tell();
        }, py::arg("position"), py::arg("whence") = py::none())
        .def("seekable", [](File&) { return true; })
        .def("tell", &File::tell)
        .def("set_len", [](File& f, size_t size) {
            std::fflush(f.f);
            if (ftruncate(fileno(f.f), (off_t)size) != 0) { PyErr_SetFromErrno(PyExc_OSError); throw py::error_already_set(); }
        }, py::arg("size"))
        .def("truncate", [](File& f) {
            std::fflush(f.f);
            if (ftruncate(fileno(f.f), 0) != 0) { PyErr_SetFromErrno(PyExc_OSError); throw py::error_already_set(); }
        })
        .def("len", &File::len)
        .def("__len__", &File::len)
        .def("__bool__", [](File& f) { return f.len() > 0; })
        .def("__repr__", [](File& f) { return "cramjam.File<path=" + f.path + ", len=" + std::to_string(f.len()) + ">"; });

    // ------------------------------------------------------------------ snappy (src/snappy.rs)
    {
        auto s = m.def_submodule("snappy", "snappy de/compression interface");
        s.def("compress", [](py::handle data, py::object output_len) { return generic_compress(CJ_SNAPPY_FRAMED, data, -1, output_len); }, py::arg("data"), py::arg("output_len") = py::none());
        s.def("decompress", [](py::handle data, py::object output_len) { return generic_decompress(CJ_SNAPPY_FRAMED, data, output_len); }, py::arg("data"),
              py::arg("output_len") = py::none());
        s.def("compress_into", [](py::handle input, py::handle output) { return generic_compress_into(CJ_SNAPPY_FRAMED, input, output, -1); }, py::arg("input"), py::arg("output"));
        s.def("decompress_into", [](py::handle input, py::handle output) { return generic_decompress_into(CJ_SNAPPY_FRAMED, input, output); }, py::arg("input"), py::arg("output"));
        // raw: output_len is accepted and ignored (src/snappy.rs:53-54,71-72)
        s.def("compress_raw", [](py::handle data, py::object) { return generic_compress(CJ_SNAPPY_RAW, data, -1); }, py::arg("data"), py::arg("output_len") = py::none());
        s.def("decompress_raw", [](py::handle data, py::object) { return generic_decompress(CJ_SNAPPY_RAW, data, py::none()); }, py::arg("data"), py::arg("output_len") = py::none());
        s.def("compress_raw_into", [](py::handle input, py::handle output) {
            size_t dw = 0;
            if (device_into(CJ_SNAPPY_RAW, true, input, output, -1, 1, &dw)) return dw;   // device arrays on both sides: CJ_DEVICE
            Input in(input);
            PyBuf out(output);  // as_bytes_mut(): slice semantics
            if ((size_t)out.b.len < cj_compress_bound(CJ_SNAPPY_RAW, in.n)) raise(g_compression_error, "snappy: output buffer (size = " + std::to_string(out.b.len) + ") is smaller than required (size = " + std::to_string(cj_compress_bound(CJ_SNAPPY_RAW, in.n)) + ")");
            Bytes c = do_compress(CJ_SNAPPY_RAW, in.p, in.n);
            std::memcpy(out.b.buf, c.data(), c.size());
            return c.size();
        }, py::arg("input"), py::arg("output"));
        s.def("decompress_raw_into", [](py::handle input, py::handle output) {
            size_t dw = 0;
            if (device_into(CJ_SNAPPY_RAW, false, input, output, -1, 1, &dw)) return dw;
            Input in(input);
            PyBuf out(output);
            size_t produced = 0;
            Bytes d = do_decompress(CJ_SNAPPY_RAW, in.p, in.n, (size_t)out.b.len, &produced);
            if (produced) std::memcpy(out.b.buf, d.data(), produced);
            return produced;
        }, py::arg("input"), py::arg("output"));
        s.def("decompress_raw_batch", [](py::iterable inputs) { return generic_batch(CJ_SNAPPY_RAW, false, inputs, -1); }, py::arg("inputs"),
              "Decompress a list of independent raw snappy blocks in one engine call; returns a list of Buffers");
        s.def("compress_raw_batch", [](py::iterable inputs) { return generic_batch(CJ_SNAPPY_RAW, true, inputs, -1); }, py::arg("inputs"),
              "Compress a list of independent buffers into raw snappy blocks in one engine call; returns a list of Buffers");
        s.def("decompress_batch", [](py::iterable inputs) { return generic_batch(CJ_SNAPPY_FRAMED, false, inputs, -1); }, py::arg("inputs"));
        s.def("compress_batch", [](py::iterable inputs) { return generic_batch(CJ_SNAPPY_FRAMED, true, inputs, -1); }, py::arg("inputs"));
        s.def("compress_raw_max_len", [](py::handle data) { Input in(data); return cj_compress_bound(CJ_SNAPPY_RAW, in.n); }, py::arg("data"));
        s.def("decompress_raw_len", [](py::handle data) {
            Input in(data);
            size_t out = 0;
            if (cj_decompressed_len(CJ_SNAPPY_RAW, in.p, in.n, &out) != 0) raise(g_decompression_error, cj_last_error());
            return out;
        }, py::arg("data"));
        add_stream_classes<0>(s, CJ_SNAPPY_FRAMED, 0, -1);
    }

    // ------------------------------------------------------------------ lz4 (src/lz4.rs)
    {
        auto l = m.def_submodule("lz4", "LZ4 de/compression interface");
        auto lvl = [](const py::object& o) { return o.is_none() ? 4 : o.cast<int>(); };  // DEFAULT_COMPRESSION_LEVEL = 4 (src/lz4.rs:17)
        l.def("compress", [lvl](py::handle data, py::object level, py::object output_len) { return generic_compress(CJ_LZ4_FRAME, data, lvl(level), output_len); }, py::arg("data"),
              py::arg("level") = py::none(), py::arg("output_len") = py::none());
        l.def("decompress", [](py::handle data, py::object output_len) { return generic_decompress(CJ_LZ4_FRAME, data, output_len); }, py::arg("data"),
              py::arg("output_len") = py::none());
        l.def("compress_into", [lvl](py::handle input, py::handle output, py::object level) { return generic_compress_into(CJ_LZ4_FRAME, input, output, lvl(level)); },
              py::arg("input"), py::arg("output"), py::arg("level") = py::none());
        l.def("decompress_into", [](py::handle input, py::handle output) { return generic_decompress_into(CJ_LZ4_FRAME, input, output); }, py::arg("input"), py::arg("output"));
        // block API; `mode` and `output_len` of compress_block are accepted and ignored (src/lz4.rs:114,120)
        l.def("compress_block", [](py::handle data, py::object, py::object, py::object acceleration, py::object compression, py::object store_size) {
            Input in(data);
            return make_buffer(lz4_block_compress(in.p, in.n, store_size.is_none() ? true : store_size.cast<bool>(), acceleration.is_none() ? 1 : acceleration.cast<int>(),
                                                  compression.is_none() ? -1 : compression.cast<int>()));
        }, py::arg("data"), py::arg("output_len") = py::none(), py::arg("mode") = py::none(), py::arg("acceleration") = py::none(),
              py::arg("compression") = py::none(), py::arg("store_size") = py::none());
        l.def("decompress_block", [](py::handle data, py::object output_len) {
            Input in(data);
            std::string err;
            size_t written = 0;
            Bytes buf;
            bool ok;
            if (!output_len.is_none()) {  // Some(n): buf = vec![0; n]; size not prepended; the full n-byte buffer is returned
                buf.assign(output_len.cast<size_t>(), 0);
                ok = lz4_block_decompress_into(in.p, in.n, buf.data(), buf.size(), false, &written, &err);
            } else {  // decompress_vec: size prefix, exact-size Vec
                if (in.n < 4) raise(g_decompression_error, "Source buffer must at least contain size prefix.");
                int32_t s;
                std::memcpy(&s, in.p, 4);
                if (s < 0) raise(g_decompression_error, "Parsed size prefix in buffer must not be negative.");
                if ((uint32_t)s > 0x7E000000u) raise(g_decompression_error, "Given size parameter is too big");
                buf.assign((size_t)s, 0);
                ok = lz4_block_decompress_into(in.p, in.n, buf.data(), buf.size(), true, &written, &err);
                if (ok) buf.resize(written);
            }
            if (!ok) raise(g_decompression_error, err);
            return make_buffer(std::move(buf));
        }, py::arg("data"), py::arg("output_len") = py::none());
        l.def("decompress_block_into", [](py::handle input, py::handle output, py::object output_len) {
            {   // device arrays on both sides: a raw block (no size prefix: the output array's size is the capacity), CJ_DEVICE
                size_t dw = 0;
                if (device_into(CJ_LZ4_BLOCK, false, input, output, -1, 1, &dw)) return dw;
            }
            Input in(input);
            const bool size_stored = output_len.is_none();
            PyBuf out(output);
            if (!output_len.is_none()) {
                const size_t size = output_len.cast<size_t>();
                if ((size_t)out.b.len < size)
                    raise(g_decompression_error, "output_len set to " + std::to_string(size) + ", but output is less. (" + std::to_string(out.b.len) + ")");
            }
            std::string e1, e2;
            size_t written = 0;
            Bytes tmp((size_t)out.b.len);
            // first the caller's stated layout, then the opposite one; the first error is the one reported (src/lz4.rs:163-170)
            bool ok = lz4_block_decompress_into(in.p, in.n, tmp.data(), tmp.size(), size_stored, &written, &e1);
            if (!ok) ok = lz4_block_decompress_into(in.p, in.n, tmp.data(), tmp.size(), !size_stored, &written, &e2);
            if (!ok) raise(g_decompression_error, e1);
            if (written) std::memcpy(out.b.buf, tmp.data(), written);
            return written;
        }, py::arg("input"), py::arg("output"), py::arg("output_len") = py::none());
        l.def("compress_block_into", [](py::handle data, py::handle output, py::object, py::object acceleration, py::object compression, py::object store_size) {
            {   // device arrays on both sides: the raw block without a size prefix, CJ_DEVICE
                size_t dw = 0;
                if (PyObject_HasAttrString(data.ptr(), "__cuda_array_interface__") && !PyObject_CheckBuffer(data.ptr()) &&
                    (store_size.is_none() || store_size.cast<bool>()))
                    raise(PyExc_ValueError, "device arrays hold raw LZ4 blocks: pass store_size=False (the size prefix is a host-side convention)");
                if (device_into(CJ_LZ4_BLOCK, true, data, output, compression.is_none() ? -1 : compression.cast<int>(),
                                acceleration.is_none() ? 1 : acceleration.cast<int>(), &dw)) return dw;
            }
            Input in(data);
            PyBuf out(output);
            Bytes c = lz4_block_compress(in.p, in.n, store_size.is_none() ? true : store_size.cast<bool>(), acceleration.is_none() ? 1 : acceleration.cast<int>(),
                                                        compression.is_none() ? -1 : compression.cast<int>());
            if (c.size() > (size_t)out.b.len) raise(g_compression_error, "Compression failed: output buffer is too small");
            std::memcpy(out.b.buf, c.data(), c.size());
            return c.size();
        }, py::arg("data"), py::arg("output"), py::arg("mode") = py::none(), py::arg("acceleration") = py::none(), py::arg("compression") = py::none(),
              py::arg("store_size") = py::none());
        l.def("decompress_block_batch", [](py::iterable inputs, py::object output_lens) {
            // output_lens=None: every block carries the 4-byte size prefix compress_block(store_size=True) writes
            std::vector<std::unique_ptr<Input>> ins;
            for (py::handle h : inputs) ins.emplace_back(new Input(h));
            std::vector<size_t> skip(ins.size(), 0), caps(ins.size());
            if (output_lens.is_none()) {
                for (size_t i = 0; i < ins.size(); i++) {
                    if (ins[i]->n < 4) raise(g_decompression_error, "unit " + std::to_string(i) + ": Source buffer must at least contain size prefix.");
                    int32_t v;
                    std::memcpy(&v, ins[i]->p, 4);
                    if (v < 0 || (uint32_t)v > 0x7E000000u) raise(g_decompression_error, "unit " + std::to_string(i) + ": bad size prefix");
                    caps[i] = (size_t)v;
                    skip[i] = 4;
                }
            } else {
                size_t i = 0;
                for (py::handle h : output_lens.cast<py::iterable>()) { if (i < caps.size()) caps[i] = h.cast<size_t>(); i++; }
                if (i != caps.size()) raise(PyExc_ValueError, "output_lens must have one entry per input");
            }
            return run_batch_call(CJ_LZ4_BLOCK, false, ins, skip, caps, -1, 1);
        }, py::arg("inputs"), py::arg("output_lens") = py::none(),
              "Decompress a list of independent LZ4 blocks in one engine call; returns a list of Buffers");
        l.def("compress_block_batch", [](py::iterable inputs, py::object acceleration, py::object compression, py::object store_size) {
            const bool prefix = store_size.is_none() ? true : store_size.cast<bool>();
            std::vector<std::unique_ptr<Input>> ins;
            for (py::handle h : inputs) ins.emplace_back(new Input(h));
            std::vector<size_t> skip(ins.size(), 0), caps(ins.size());
            for (size_t i = 0; i < ins.size(); i++) caps[i] = cj_compress_bound(CJ_LZ4_BLOCK, ins[i]->n);
            py::list raw = run_batch_call(CJ_LZ4_BLOCK, true, ins, skip, caps, compression.is_none() ? -1 : compression.cast<int>(),
                                          acceleration.is_none() ? 1 : acceleration.cast<int>());
            if (!prefix) return raw;
            py::list out;
            for (size_t i = 0; i < ins.size(); i++) {
                Buffer& b = raw[i].cast<Buffer&>();
                Bytes v(b.own.size() + 4);
                const uint32_t ul = (uint32_t)ins[i]->n;
                std::memcpy(v.data(), &ul, 4);
                if (!b.own.empty()) std::memcpy(v.data() + 4, b.own.data(), b.own.size());
                out.append(make_buffer(std::move(v)));
            }
            return out;
        }, py::arg("inputs"), py::arg("acceleration") = py::none(), py::arg("compression") = py::none(), py::arg("store_size") = py::none());
        l.def("decompress_batch", [](py::iterable inputs) { return generic_batch(CJ_LZ4_FRAME, false, inputs, -1); }, py::arg("inputs"));
        l.def("compress_batch", [lvl](py::iterable inputs, py::object level) { return generic_batch(CJ_LZ4_FRAME, true, inputs, lvl(level)); }, py::arg("inputs"),
              py::arg("level") = py::none());
        l.def("compress_block_bound", [](py::handle src) { Input in(src); return cj_compress_bound(CJ_LZ4_BLOCK, in.n) + 4; }, py::arg("src"));
        add_stream_classes<1>(l, CJ_LZ4_FRAME, 2, 4);
    }

    // ------------------------------------------------------------------ zstd (src/zstd.rs)
    {
        auto z = m.def_submodule("zstd", "zstd de/compression interface");
        auto lvl = [](const py::object& o) { return o.is_none() ? 0 : o.cast<int>(); };  // DEFAULT_COMPRESSION_LEVEL = 0 -> libzstd default (src/zstd.rs:14)
        z.def("compress", [lvl](py::handle data, py::object level, py::object output_len) { return generic_compress(CJ_ZSTD, data, lvl(level), output_len); }, py::arg("data"),
              py::arg("level") = py::none(), py::arg("output_len") = py::none());
        z.def("decompress", [](py::handle data, py::object output_len) { return generic_decompress(CJ_ZSTD, data, output_len); }, py::arg("data"),
              py::arg("output_len") = py::none());
        z.def("compress_into", [lvl](py::handle input, py::handle output, py::object level) { return generic_compress_into(CJ_ZSTD, input, output, lvl(level)); },
              py::arg("input"), py::arg("output"), py::arg("level") = py::none());
        z.def("decompress_into", [](py::handle input, py::handle output) { return generic_decompress_into(CJ_ZSTD, input, output); }, py::arg("input"), py::arg("output"));
        z.def("decompress_batch", [](py::iterable inputs) { return generic_batch(CJ_ZSTD, false, inputs, -1); }, py::arg("inputs"),
              "Decompress a list of independent zstd streams in one engine call; returns a list of Buffers");
        z.def("compress_batch", [lvl](py::iterable inputs, py::object level) { return generic_batch(CJ_ZSTD, true, inputs, lvl(level)); }, py::arg("inputs"),
              py::arg("level") = py::none());
        add_stream_classes<2>(z, CJ_ZSTD, 1, 0);
    }
}
