// oip_cli.cpp -- "OpticalImageProcessor" command line: the reference's process-level contract
// (argv grammar, cwd-relative output names, exit codes, LOGFILE) re-implemented in C++17 on top of
// the C ABI (include/oip_b200.h).  ref main.cpp:92-343 (grammar/exit codes), imageop.h:99-108 (names),
// DOC/Usage.txt (task flow).  CLI11 is not available here: hand-written parser, same option names.
//
// Differences, all forced by what SURVEY 8f leaves for later rows:
//   * prestitch: the inter-CMOS offset estimate (ref stitcher.h:148-201) runs on the GPU (oip_stt_parameters; agrees with
//     cv::phaseCorrelate to ~1e-3 px); the extension options --dx/--dy skip it.
//   * default action: the inter-band correlation + polynomial fit (ref preproc.h:224-347,492-550) runs on the GPU
//     (oip_inter_band_correlation); the extension option --poly FILE (4 lines of cx0 cx1 cy0 cy1 cy2) skips it.
//   * TIFF files (host/tiff_io.hpp): written uncompressed (the reference's libraries compress with LZW), so
//     they match in geometry, sample order and pixel values, not byte for byte; TIFF INPUT may be uncompressed or LZW
//     (predictor 1 / 2) strips, i.e. our own products and the reference's.  --aligned-raw additionally writes
//     <stem>.ALIGNED.RAW (the CV_16UC4 memory image).
// There is no CPU fallback: without a B200 every command fails with exit code 2.
#include <strings.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <filesystem>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/oip_b200.h"
#include "tiff_io.hpp"

#include <atomic>
#include <fcntl.h>
#include <future>
#include <mutex>
#include <sys/mman.h>
#include <thread>
#include <unistd.h>
namespace fs = std::filesystem;

// ---- reference constants (ref oipshared.h:27-64)
static const int PIXELS_PER_LINE = 12288, BYTES_PER_PIXEL = 2, MSS_BANDS = 4;
static const int REMAP_ROW_GUARD = 32767, REMAP_SECTION_ROWS = 30000; // ref imageop.h:19-20

struct usage_error : std::invalid_argument { using std::invalid_argument::invalid_argument; };   // ref main.cpp:20-23
struct parse_error : std::runtime_error { int code; parse_error(int c, const std::string &m) : std::runtime_error(m), code(c) {} };
enum { CLI_VALIDATION = 105, CLI_REQUIRED = 106, CLI_REQUIRES = 107, CLI_EXTRAS = 109, CLI_CONVERSION = 104 }; // CLI11 ExitCodes

// ---- logging: LOGFILE or oip.log, timestamped (ref main.cpp:319-329)
static FILE *g_log = nullptr;
static void logf(const char *lvl, const char *fmt, ...)
{
    char msg[2048];
    va_list ap; va_start(ap, fmt); vsnprintf(msg, sizeof msg, fmt, ap); va_end(ap);
    time_t t = time(nullptr); char ts[32]; strftime(ts, sizeof ts, "%Y-%m-%d %H:%M:%S", localtime(&t));
    if (g_log) { fprintf(g_log, "%s [%s] %s\n", ts, lvl, msg); fflush(g_log); }
    printf("%s\n", msg);
}
#define OLOG(...) logf("T", __VA_ARGS__)

static void oip_check(int rc)
{
    if (rc == OIP_OK) return;
    if (rc == OIP_E_INVALID) throw std::invalid_argument(oip_last_error());
    throw std::runtime_error(oip_last_error());
}
static oip_ctx *ctx()
{
    static oip_ctx *c = nullptr;
    if (!c) oip_check(oip_ctx_create(0, nullptr, 1, &c));
    return c;
}

// ---- file helpers (ref imageop.h:43-108)
static size_t file_size(const std::string &p)
{
    struct stat st {};
    if (stat(p.c_str(), &st)) throw std::runtime_error("stat() call for file failed: " + p);
    return (size_t)st.st_size;
}
static std::string build_output_path(const std::string &tmpl, const std::string &stem_ext, const char *replace_ext = nullptr)
{
    fs::path t = tmpl;                                       // ref imageop.h:99-108
    fs::path o = fs::current_path() / t.stem();
    o += stem_ext;
    o += replace_ext ? replace_ext : t.extension().string();
    return o.string();
}
// TIFF products carry the options the reference's libraries apply (LZW + predictor 2); OIP_TIFF_COMPRESS=none writes them
// uncompressed (faster on a RAM disk, pixel-identical)
static int tiff_compression()
{
    const char *e = getenv("OIP_TIFF_COMPRESS");
    return (e && (!strcmp(e, "none") || !strcmp(e, "NONE") || !strcmp(e, "0"))) ? oiptiff::COMPRESS_NONE : oiptiff::COMPRESS_LZW;
}
static std::string lower(std::string s) { std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)tolower(c); }); return s; }

struct Pinned {
    void *p = nullptr; size_t n = 0;
    explicit Pinned(size_t bytes) : n(bytes) { oip_check(oip_host_alloc_pinned(bytes ? bytes : 1, &p)); }
    ~Pinned() { if (p) oip_host_free_pinned(p); }
    Pinned(const Pinned &) = delete;
};
struct DevBuf {
    void *p = nullptr;
    explicit DevBuf(size_t bytes) { oip_check(oip_dev_alloc(ctx(), bytes ? bytes : 1, &p)); }
    ~DevBuf() { if (p) oip_dev_free(ctx(), p); }
    DevBuf(const DevBuf &) = delete;
};
static void write_file(const std::string &path, const void *src, size_t bytes)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("open file [" + path + "] failed");
    size_t done = 0;
    while (done < bytes) {
        size_t n = fwrite((const char *)src + done, 1, std::min<size_t>(8u << 20, bytes - done), f);
        if (!n) { fclose(f); throw std::runtime_error("write file failed: " + path); }
        done += n;
    }
    fclose(f);
}

// ---------------------------------------------------------------------------------------------
// Streamed file <-> device transfers (SURVEY 0.1 / north_star: host file I/O overlaps with device work).  The reference
// reads whole files into memory before it starts (ref imageop.h:52-97); round 1 of this CLI did the same into one pinned
// buffer.  Here a file moves in 256 MiB chunks through two pinned buffers: while chunk k travels over PCIe
// (cudaMemcpyAsync on the context's stream) the I/O threads already read chunk k+1 (pread on disjoint slices), and
// the other way round for products (pwrite of chunk k while chunk k+1 comes down).
// ---------------------------------------------------------------------------------------------
static unsigned io_threads()
{
    unsigned n = std::thread::hardware_concurrency();
    return std::max(1u, std::min(8u, n ? n : 1u));
}
struct Fd {
    int fd = -1;
    Fd(const std::string &path, int flags, const char *err) { fd = ::open(path.c_str(), flags, 0644); if (fd < 0) throw std::invalid_argument(std::string(err) + " [" + path + "]"); }
    ~Fd() { if (fd >= 0) ::close(fd); }
    Fd(const Fd &) = delete;
};
// all of [off, off + bytes) of the file <-> buf, split over the I/O threads
static void parallel_io(bool wr, int fd, void *buf, size_t bytes, size_t off, const std::string &path)
{
    const unsigned nt = (unsigned)std::min<size_t>(io_threads(), std::max<size_t>(1, bytes >> 22));
    std::vector<std::thread> pool;
    std::atomic<bool> ok{true};
    const size_t per = ((bytes + nt - 1) / nt + 4095) & ~(size_t)4095;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([&, t]() {
            size_t a = std::min(bytes, (size_t)t * per), b = std::min(bytes, a + per);
            while (a < b) {
                const ssize_t n = wr ? ::pwrite(fd, (const char *)buf + a, b - a, (off_t)(off + a)) : ::pread(fd, (char *)buf + a, b - a, (off_t)(off + a));
                if (n <= 0) { ok = false; return; }
                a += (size_t)n;
            }
        });
    for (auto &th : pool) th.join();
    if (!ok) throw std::runtime_error(wr ? "write file failed: " + path : "file size doesn't match with read byte count: " + path);
}
static const size_t IO_CHUNK = 256u << 20;
// file bytes [offset, offset + bytes) -> device memory
static void upload_file(const std::string &path, size_t offset, size_t bytes, void *d_dst)
{
    Fd f(path, O_RDONLY, "cannot open file");
    if (!bytes) return;
    const size_t ch = std::min(bytes, IO_CHUNK);
    Pinned a(ch), b(bytes > ch ? ch : 1);
    void *buf[2] = {a.p, b.p};
    parallel_io(false, f.fd, buf[0], ch, offset, path);
    int cur = 0;
    for (size_t done = 0; done < bytes;) {
        const size_t n = std::min(ch, bytes - done);
        oip_check(oip_copy_h2d(ctx(), (char *)d_dst + done, buf[cur], n));          // async: overlaps the next read
        const size_t next = done + n;
        if (next < bytes) parallel_io(false, f.fd, buf[cur ^ 1], std::min(ch, bytes - next), offset + next, path);
        oip_check(oip_ctx_sync(ctx()));
        done = next;
        cur ^= 1;
    }
}
// device memory -> file bytes [offset, offset + bytes) (the file is created / truncated when offset == 0)
static void download_to_file(const std::string &path, const void *d_src, size_t bytes, size_t offset = 0)
{
    Fd f(path, O_WRONLY | O_CREAT | (offset ? 0 : O_TRUNC), "open file failed:");
    if (!bytes) return;
    const size_t ch = std::min(bytes, IO_CHUNK);
    Pinned a(ch), b(bytes > ch ? ch : 1);
    void *buf[2] = {a.p, b.p};
    oip_check(oip_copy_d2h(ctx(), buf[0], d_src, std::min(ch, bytes)));
    oip_check(oip_ctx_sync(ctx()));
    int cur = 0;
    for (size_t done = 0; done < bytes;) {
        const size_t n = std::min(ch, bytes - done), next = done + n;
        if (next < bytes) oip_check(oip_copy_d2h(ctx(), buf[cur ^ 1], (const char *)d_src + next, std::min(ch, bytes - next))); // async
        parallel_io(true, f.fd, buf[cur], n, offset + done, path);                   // overlaps the next copy
        oip_check(oip_ctx_sync(ctx()));
        done = next;
        cur ^= 1;
    }
}
static std::vector<double> load_rrc(const std::string &path, int cols)
{
    std::vector<double> kb((size_t)cols * 2);
    oip_check(oip_load_rrc_csv(path.c_str(), cols, kb.data()));
    return kb;
}

// ---- tiny option parser with CLI11's conventions (--opt=v, --opt v, -o v, flags, positionals)
struct Args {
    std::map<std::string, std::string> val;
    std::vector<std::string> pos;
    bool has(const std::string &k) const { return val.count(k) != 0; }
    std::string get(const std::string &k, const std::string &d = "") const { auto i = val.find(k); return i == val.end() ? d : i->second; }
    long long geti(const std::string &k, long long d) const {
        if (!has(k)) return d;
        char *e = nullptr; long long v = strtoll(get(k).c_str(), &e, 10);
        if (!e || *e) throw parse_error(CLI_CONVERSION, "Could not convert: --" + k + " = " + get(k));
        return v;
    }
    double getd(const std::string &k, double d) const {
        if (!has(k)) return d;
        char *e = nullptr; double v = strtod(get(k).c_str(), &e);
        if (!e || *e) throw parse_error(CLI_CONVERSION, "Could not convert: --" + k + " = " + get(k));
        return v;
    }
};
struct OptSpec { const char *name; const char *alias; bool flag; };
static Args parse(const std::vector<std::string> &argv, const std::vector<OptSpec> &specs, int max_pos)
{
    Args a;
    auto find = [&](const std::string &n) -> const OptSpec * {
        for (auto &s : specs) if (n == s.name || (s.alias && n == s.alias)) return &s;
        return nullptr;
    };
    for (size_t i = 0; i < argv.size(); ++i) {
        const std::string &t = argv[i];
        if (t.size() > 1 && t[0] == '-') {
            std::string key = t, v; bool hasv = false;
            size_t eq = t.find('=');
            if (eq != std::string::npos) { key = t.substr(0, eq); v = t.substr(eq + 1); hasv = true; }
            const OptSpec *s = find(key);
            if (!s) throw parse_error(CLI_EXTRAS, "The following argument was not expected: " + t);
            std::string canon = std::string(s->name).substr(2);
            if (s->flag) { a.val[canon] = "1"; continue; }
            if (!hasv) {
                if (i + 1 >= argv.size()) throw parse_error(114, key + ": 1 required");
                v = argv[++i];
            }
            a.val[canon] = v;
        } else {
            if ((int)a.pos.size() >= max_pos) throw parse_error(CLI_EXTRAS, "The following argument was not expected: " + t);
            a.pos.push_back(t);
        }
    }
    return a;
}
static void require(const Args &a, const char *k) { if (!a.has(k)) throw parse_error(CLI_REQUIRED, std::string("--") + k + " is required"); }
static void existing_file(const std::string &p, const char *what)
{
    struct stat st {};
    if (stat(p.c_str(), &st) || !S_ISREG(st.st_mode)) throw parse_error(CLI_VALIDATION, std::string(what) + ": File does not exist: " + p);
}

// =============================================================================================
// auxsep  (ref main.cpp:98-109, aux_separator.h:193-327)
// =============================================================================================
struct AosFileInfo { char station[16], satellite[16]; int year, month, day, hour, minute, second; };
static bool parse_file_info(const char *name, AosFileInfo &afi)     // ref aux_separator.h:692-719 (same sscanf pattern)
{
    int cmos = 0; char y[5], mo[3], d[3], h[3], mi[3], s[3];
    if (sscanf(name, "%15[A-Za-z0-9]%*[_-]%15[A-Za-z0-9-]_%4[0-9]%2[0-9]%2[0-9]_%2[0-9]%2[0-9]%2[0-9]_%d", afi.station, afi.satellite,
               y, mo, d, h, mi, s, &cmos) != 9) return false;
    afi.year = atoi(y); afi.month = atoi(mo); afi.day = atoi(d); afi.hour = atoi(h); afi.minute = atoi(mi); afi.second = atoi(s);
    return true;
}

static int cmd_auxsep(const std::vector<std::string> &av)
{
    Args a = parse(av, {{"--offset", "-O", false}}, 1);
    if (a.pos.empty()) throw parse_error(CLI_REQUIRED, "file is required");
    const std::string file = a.pos[0];
    existing_file(file, "file");
    size_t offset = (size_t)a.geti("offset", 0);
    const size_t ps = (size_t)getpagesize();
    if (offset % ps) { offset = offset / ps * ps; OLOG("offset not aligned with system memory page size, adjusted to %zu (0x%zX).", offset, offset); }
    const bool is_imdt = strcasecmp(fs::path(file).extension().string().c_str(), ".IMDT") == 0;   // ref :204-206

    std::string imdt_name = file;
    std::vector<uint8_t> keep_host;
    DevBuf *d_imdt = nullptr;
    int64_t imdt_bytes = 0;
    std::unique_ptr<DevBuf> d_file, d_payload, d_imdt_own;
    if (!is_imdt) {
        AosFileInfo afi{};
        if (!parse_file_info(fs::path(file).filename().string().c_str(), afi) &&
            !parse_file_info(fs::path(file).parent_path().filename().string().c_str(), afi))
            throw std::invalid_argument("unrecognized AOS file name pattern");                       // ref :208-213
        OLOG("Launching AOS file separation ...");
        const size_t total = file_size(file);
        if (offset > total) throw std::invalid_argument("offset beyond end of file");
        const size_t n = total - offset;
        d_file.reset(new DevBuf(n));
        upload_file(file, offset, n, d_file->p);
        const size_t cap = n / 1024 + 1;
        d_payload.reset(new DevBuf(cap * 8));
        int64_t cnt[3];
        auto t0 = std::chrono::steady_clock::now();
        oip_check(oip_aos_scan(ctx(), (const uint8_t *)d_file->p, n, (uint64_t *)d_payload->p, cap, cnt));
        OLOG("%lld valid, %lld invalid, %lld empty AOS frames.", (long long)cnt[0], (long long)cnt[1], (long long)cnt[2]);
        const size_t icap = (size_t)(cnt[0] * 880 / 882 + 1) * 866;
        d_imdt_own.reset(new DevBuf(icap));
        int64_t st[9];
        oip_check(oip_imtr_deframe(ctx(), (const uint8_t *)d_file->p, (const uint64_t *)d_payload->p, cnt[0], (uint8_t *)d_imdt_own->p,
                                   icap, st, &imdt_bytes));
        double es = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        OLOG("%zu bytes processed for AOS filemap in %.3f seconds (%.1f MBps).", n, es, n / es / (1024.0 * 1024.0));
        OLOG("%lld image transfer frames cut, %lld accepted (bad sig %lld, tail %lld, type %lld, CRC %lld).", (long long)st[0],
             (long long)st[1], (long long)st[2], (long long)st[3], (long long)st[4], (long long)st[5]);
        if (st[1] == 0) { OLOG("No more AOS frame data, end of job."); return 0; }
        char nm[256];
        snprintf(nm, sizeof nm, "%s_%s_%s_%04d%02d%02d_%02d%02d%02d.IMDT", afi.station, afi.satellite,
                 st[7] == 0x11 ? "CMOS-1" : "CMOS-2", afi.year, afi.month, afi.day, afi.hour, afi.minute, afi.second); // ref :514-523
        imdt_name = nm;
        download_to_file(imdt_name, d_imdt_own->p, (size_t)imdt_bytes);                               // cwd, ref :524
        d_imdt = d_imdt_own.get();
        OLOG("Parsing done.");
    } else {
        imdt_bytes = (int64_t)file_size(file);
        d_imdt_own.reset(new DevBuf((size_t)imdt_bytes));
        upload_file(file, 0, (size_t)imdt_bytes, d_imdt_own->p);
        d_imdt = d_imdt_own.get();
    }
    OLOG("Separating aux & image data ...");
    const oip_frame_geom g{1536, 256};                                                                // ref :92-93
    int64_t fst[4];
    // ONE index pass: the table holds every complete frame the stream can contain (+ slack for zero-filled gap frames);
    // only a stream with long sequence gaps needs the second call
    int64_t cap_fr = imdt_bytes / (49152 + 40ll * 1536 * 256 * 2 + 172) + 64;
    std::vector<oip_frame_entry> ents((size_t)cap_fr);
    int irc = oip_image_frames_index(ctx(), (const uint8_t *)d_imdt->p, (size_t)imdt_bytes, &g, ents.data(), cap_fr, fst);
    if (irc == OIP_E_INVALID && fst[1] > cap_fr) {
        cap_fr = fst[1];
        ents.resize((size_t)cap_fr);
        irc = oip_image_frames_index(ctx(), (const uint8_t *)d_imdt->p, (size_t)imdt_bytes, &g, ents.data(), cap_fr, fst);
    }
    oip_check(irc);
    const int64_t nf = fst[1];
    const size_t aux_b = (size_t)nf * 49152, pan_b = (size_t)nf * 1024 * 12288 * 2, mss_b = (size_t)nf * 256 * 12288 * 2;
    const std::string aux_name = build_output_path(imdt_name, "", ".AUX");                            // ref :260-262
    const std::string pan_name = build_output_path(imdt_name, ".PAN", ".RAW");
    const std::string mss_name = build_output_path(imdt_name, ".MSS", ".RAW");
    if (nf > 0) {
        DevBuf d_aux(aux_b), d_pan(pan_b), d_mss(mss_b);
        oip_check(oip_unpack_frames(ctx(), (const uint8_t *)d_imdt->p, (size_t)imdt_bytes, &g, ents.data(), nf, (uint8_t *)d_aux.p,
                                    (uint16_t *)d_pan.p, (uint16_t *)d_mss.p));
        oip_check(oip_ctx_sync(ctx()));
        download_to_file(aux_name, d_aux.p, aux_b);
        download_to_file(pan_name, d_pan.p, pan_b);
        download_to_file(mss_name, d_mss.p, mss_b);
    } else {
        write_file(aux_name, "", 0); write_file(pan_name, "", 0); write_file(mss_name, "", 0);        // the reference creates them empty
    }
    OLOG("%4lld image frames processed.", (long long)fst[3]);
    OLOG("Done.");
    return 0;
}

// =============================================================================================
// prestitch  (ref main.cpp:112-150, :270-286; stitcher.h)
// =============================================================================================
static int cmd_prestitch(const std::vector<std::string> &av)
{
    Args a = parse(av, {{"--pan1", nullptr, false}, {"--pan2", nullptr, false}, {"--rrc1", nullptr, false}, {"--rrc2", nullptr, false},
                        {"--sections", "-s", false}, {"--section-lines", "-l", false}, {"--stitch-overlap", nullptr, false},
                        {"--stt-threshold", nullptr, false}, {"--stt-maxdeltay", nullptr, false}, {"--edge-cols", "-e", false},
                        {"--rrc", "-r", true}, {"--no-rrc", nullptr, true}, {"--only-calculate", "-c", true},
                        {"--dx", nullptr, false}, {"--dy", nullptr, false}}, 0);
    require(a, "pan1"); require(a, "pan2");
    const std::string pan1 = a.get("pan1"), pan2 = a.get("pan2");
    existing_file(pan1, "--pan1"); existing_file(pan2, "--pan2");
    if (a.has("rrc1")) existing_file(a.get("rrc1"), "--rrc1");
    if (a.has("rrc2")) existing_file(a.get("rrc2"), "--rrc2");
    const long long overlap = a.geti("stitch-overlap", 200), edge = a.geti("edge-cols", 0);
    if (edge < 0 || edge > overlap / 2) throw parse_error(CLI_VALIDATION, "--edge-cols: invalid edge cols");   // ref main.cpp:135-141
    const int sections = (int)a.geti("sections", 10), sec_lines = (int)a.geti("section-lines", 16000);
    const bool do_rrc = !a.has("no-rrc");
    // Stitcher::Stitcher size checks, ref stitcher.h:60-77
    const size_t s1 = file_size(pan1), s2 = file_size(pan2);
    if ((size_t)sections * sec_lines * BYTES_PER_PIXEL > s1) throw std::invalid_argument("PAN1 size too small for SECTION & LINE_PER_SECTION argument");
    if ((size_t)sections * sec_lines * BYTES_PER_PIXEL > s2) throw std::invalid_argument("PAN2 size too small for SECTION & LINE_PER_SECTION argument");
    if (s1 != s2) throw std::invalid_argument("PAN1 size doesn't match PAN2 size");
    const int64_t lines = (int64_t)(s1 / (PIXELS_PER_LINE * BYTES_PER_PIXEL));
    OLOG("PAN: %lld lines total.", (long long)lines);
    if (lines < (int64_t)sections * sec_lines)
        throw std::invalid_argument("PAN line count less than sections times line-per-section, use smaller -s and/or -l value(s)");
    double dx = 0.0, dy = 0.0;
    if (a.has("dx") && a.has("dy")) { // extension: skip the estimate
        dx = a.getd("dx", 0); dy = a.getd("dy", 0);
        OLOG("    dx: %.5f, dy: %.5f (given)", dx, dy);
    } else {
        // Stitcher::CalcSttParameters, ref stitcher.h:148-201.  It runs on the files as given: PreStitch() calls it
        // before DoRRC (ref main.cpp:280-284) and mRrcFilePAN1/2 still name the inputs (stitcher.h:79-80).
        const size_t nb = s1;
        DevBuf d1(nb), d2(nb);
        upload_file(pan1, 0, nb, d1.p);
        upload_file(pan2, 0, nb, d2.p);
        oip_stt_config cfg{};
        cfg.sections = sections; cfg.lines_per_section = sec_lines; cfg.overlap_cols = (int)overlap; cfg.edge_cols = (int)edge;
        cfg.threshold = a.getd("stt-threshold", 0.4); cfg.max_delta_y = a.getd("stt-maxdeltay", 0.0);
        std::vector<oip_stt_section> secs((size_t)sections);
        double sums[4];
        OLOG("Calculating stitching delta values ...");
        oip_check(oip_stt_parameters(ctx(), (const uint16_t *)d1.p, (const uint16_t *)d2.p, PIXELS_PER_LINE, lines, 0, lines,
                                     PIXELS_PER_LINE, &cfg, secs.data(), sums));
        OLOG("| offset |  delta x |  delta y | response | r |");
        OLOG("-----------------------------------------------");
        for (const oip_stt_section &q : secs)
            OLOG("|%7lld |%10.4f|%10.4f|%10.4f|%s|", (long long)q.line_offset, q.dx, q.dy, q.response, q.valid == 1 ? " + " : " x ");
        if (sums[3] == 0.0) throw std::runtime_error("No valid delta value found for stitching parameter calculating");
        dx = sums[0] / sums[3]; dy = sums[1] / sums[3];
        OLOG("Total %d valid delta value pairs found, everage value:", (int)sums[3]);
        OLOG("    dx: %.5f, dy: %.5f, r: %.5f", dx, dy, sums[2] / sums[3]);
    }
    if (a.has("only-calculate")) return 0;
    if (do_rrc && (!a.has("rrc1") || !a.has("rrc2"))) throw std::runtime_error("open RRC Param file failed");

    const size_t bytes = s1;
    DevBuf d_a(bytes), d_b(bytes);
    std::string rrc2_path = pan2;
    if (do_rrc) {                                                                                     // Stitcher::DoRRC, ref stitcher.h:141-146
        const std::string p1 = build_output_path(pan1, ".RRC"), p2 = build_output_path(pan2, ".RRC");
        for (int i = 0; i < 2; ++i) {
            const std::string &src = i ? pan2 : pan1;
            std::vector<double> kb = load_rrc(a.get(i ? "rrc2" : "rrc1"), PIXELS_PER_LINE);
            DevBuf d_kb(kb.size() * 8);
            upload_file(src, 0, bytes, d_a.p);
            oip_check(oip_copy_h2d(ctx(), d_kb.p, kb.data(), kb.size() * 8));
            OLOG("Do inplace RRC ...");
            oip_check(oip_rrc_u16(ctx(), (uint16_t *)d_a.p, PIXELS_PER_LINE, lines, PIXELS_PER_LINE, (const double *)d_kb.p));
            oip_check(oip_ctx_sync(ctx()));
            OLOG("Write RRC result as file \"%s\" ...", (i ? p2 : p1).c_str());
            download_to_file(i ? p2 : p1, d_a.p, bytes);
        }
        rrc2_path = p2; // d_a now holds the corrected PAN2
    } else {
        upload_file(pan2, 0, bytes, d_a.p);
    }
    // Stitcher::PreStitch, ref stitcher.h:83-139
    if (lines <= REMAP_ROW_GUARD) throw std::invalid_argument("too few data rows, please use cv::remap()");   // ref imageop.h:242-244
    const std::string out = build_output_path(rrc2_path, ".PRESTT");
    oip_check(oip_shift_cubic_u16(ctx(), (const uint16_t *)d_a.p, (uint16_t *)d_b.p, PIXELS_PER_LINE, lines, dx, dy, REMAP_SECTION_ROWS,
                                  REMAP_ROW_GUARD));
    oip_check(oip_pan_check_error(ctx()));
    download_to_file(out, d_b.p, bytes);
    OLOG("Pre-stitched PAN2 written to file '%s'.", out.c_str());
    return 0;
}

// =============================================================================================
// stitch  (ref main.cpp:153-190, stitcher.h:21-46, imageop.h:277-363)
// =============================================================================================
static int cmd_stitch(const std::vector<std::string> &av)
{
    Args a = parse(av, {{"--image1", nullptr, false}, {"--image2", nullptr, false}, {"--out", "-o", false}, {"--fold-cols", "-c", false},
                        {"--GDAL", "-g", true}, {"--band-map", "-m", false}}, 0);
    require(a, "image1"); require(a, "image2"); require(a, "fold-cols");
    const int fold = (int)a.geti("fold-cols", 0);
    if (fold < 2) throw parse_error(CLI_VALIDATION, "--fold-cols: fold column value too small");              // ref main.cpp:166-170
    if (a.has("band-map") && !a.has("GDAL")) throw parse_error(CLI_REQUIRES, "--band-map requires --GDAL");     // ref :175-176
    if (a.has("band-map")) {
        int m[4];
        if (sscanf(a.get("band-map").c_str(), "%d,%d,%d,%d", m, m + 1, m + 2, m + 3) != 4) throw parse_error(CLI_VALIDATION, "-m: need 4 band indices");
        for (int i = 0; i < 4; ++i) if (m[i] <= 0 || m[i] > MSS_BANDS) throw parse_error(CLI_VALIDATION, "-m: invalid band index");
    }
    const std::string l = a.get("image1"), r = a.get("image2"), out = a.get("out");
    const std::string le = lower(fs::path(l).extension().string()), re = lower(fs::path(r).extension().string());
    if (le != re) throw std::invalid_argument("Stitch(): two images should be same type");                    // ref stitcher.h:31-33
    if (le != ".tiff" && le != ".raw") throw std::invalid_argument("Stitch(): only RAW and TIFF image supported");
    if (le == ".tiff") { // IMO::StitchTiff / StitchTiffGDAL, ref imageop.h:365-567
        std::string outp = out;
        if (outp.empty()) outp = (fs::current_path() / "stitched.TIFF").string();                             // ref :373-375
        else if (lower(fs::path(outp).extension().string()) != ".tiff") throw std::invalid_argument("Output file should be a tiff image"); // :376-380
        OLOG("Reading tiff image from file `%s' ...", l.c_str());
        const oiptiff::Info il = oiptiff::read_info(l), ir = oiptiff::read_info(r);
        OLOG("Image size: %lld cols, %lld rows, channels: %d.", (long long)il.width, (long long)il.height, il.spp);
        if (il.height != ir.height || il.width != ir.width) throw std::runtime_error("images have different sizes");   // :412-414
        if (il.spp != 4 || ir.spp != 4) throw std::runtime_error("4-channel 16-bit TIFF images expected");
        const int f = fold / 2, w = (int)il.width, ow = 2 * (w - f);
        if (f >= w) throw std::invalid_argument("fold columns exceed the image width");
        const size_t nb = (size_t)il.height * w * 8, ob = (size_t)il.height * ow * 8;
        Pinned hl(nb), hr(nb), ho(ob);
        oiptiff::read_u16(l, il, (uint16_t *)hl.p);
        oiptiff::read_u16(r, ir, (uint16_t *)hr.p);
        DevBuf dl(nb), dr(nb), dout(ob);
        oip_check(oip_copy_h2d(ctx(), dl.p, hl.p, nb));
        oip_check(oip_copy_h2d(ctx(), dr.p, hr.p, nb));
        // Samples travel in FILE order.  cv::imwrite / cv::imread swap samples 0 and 2 of a 4-channel image (BGRA memory,
        // RGBA file), so the cv path (ref :419-446) keeps the file order; the GDAL path (ref :522-537) writes band b from
        // MEMORY channel map[b]-1, i.e. from file sample swap02(map[b]-1).
        const bool gdal = a.has("GDAL") || file_size(l) >= 4000000000ull;                                      // ref :418
        int map[4] = {1, 2, 3, 4};
        if (a.has("band-map")) sscanf(a.get("band-map").c_str(), "%d,%d,%d,%d", map, map + 1, map + 2, map + 3);
        int fmap[4];
        for (int b = 0; b < 4; ++b) { const int m = map[b] - 1; fmap[b] = (m == 0 ? 2 : (m == 2 ? 0 : m)) + 1; }
        const uint16_t *imgs[2] = {(const uint16_t *)dl.p, (const uint16_t *)dr.p};
        OLOG("Begin stitching two images ...");
        oip_check(oip_stitch_concat_c4(ctx(), imgs, 2, w, il.height, f, gdal ? fmap : nullptr, (uint16_t *)dout.p));
        oip_check(oip_copy_d2h(ctx(), ho.p, dout.p, ob));
        oip_check(oip_ctx_sync(ctx()));
        OLOG("Write stitched image to file '%s' ...", outp.c_str());
        // cv::imwrite's TIFF default and the GTiff options of ref imageop.h:470-474: LZW + horizontal predictor; OIP_TIFF_COMPRESS=none opts out
        oiptiff::write_u16(outp, (const uint16_t *)ho.p, ow, il.height, 4, 2, tiff_compression());
        OLOG("%zu bytes written.", ob);
        return 0;
    }
    std::string raw_out = out;
    bool out_tiff = true;                                                                                       // ref imageop.h:297-306
    if (out.empty()) raw_out = (fs::current_path() / ("stitched_" + std::to_string(2 * (PIXELS_PER_LINE - fold / 2)) + "n16b.TIFF")).string();
    else out_tiff = lower(fs::path(out).extension().string()) == ".tiff";
    const size_t szl = file_size(l), szr = file_size(r);
    if (szl != szr) throw std::invalid_argument("RAW image sizes not match");                                   // ref imageop.h:285-289
    const int f = fold / 2;                                                                                     // ref main.cpp:189
    const int64_t lines = (int64_t)(szl / (PIXELS_PER_LINE * BYTES_PER_PIXEL));
    const int out_w = oip_pan_out_width(2, PIXELS_PER_LINE, f);
    // Streamed: the strip is cut into row windows; several workers (own context = own streams) each read a window of both
    // files into pinned memory, run it through the host-buffer pipeline (H2D / fused kernel / D2H overlapped in row blocks
    // inside oip_pan_pipeline_host) and write its output rows at their place in the product -- so file reads, both PCIe
    // directions and file writes of different windows overlap.  (The reference reads line by line, ref imageop.h:330-357.)
    const size_t out_bytes = (size_t)lines * out_w * 2;
    size_t data0 = 0;
    if (out_tiff) data0 = (size_t)oiptiff::write_u16(raw_out, nullptr, out_w, lines, 1, 1);   // single-band GTiff, ref imageop.h:316-328: header first
    // the product is written through a shared mapping: concurrent pwrite()s to ONE file serialise on the inode lock (measured
    // on tmpfs: 4 and 16 workers took the same 4.7 s for 6.4 GB), stores into a mapping do not
    Fd fo_map(raw_out, O_RDWR | O_CREAT | (out_tiff ? 0 : O_TRUNC), "open file failed:");
    if (ftruncate(fo_map.fd, (off_t)(data0 + out_bytes))) throw std::runtime_error("write file failed: " + raw_out);
    // (the file's pages are allocated by the page faults of the parallel stores below: 8 threads fault ~4x faster than one
    // posix_fallocate call allocates -- measured 0.4 s against 2 s for 6.4 GB on this box's RAM disk)
    uint8_t *out_map = (uint8_t *)mmap(nullptr, data0 + out_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fo_map.fd, 0);
    if (out_map == MAP_FAILED) throw std::runtime_error("mmap of the output file failed: " + raw_out);
    struct Unmap { void *p; size_t n; ~Unmap() { munmap(p, n); } } unmap{out_map, data0 + out_bytes};
    OLOG("Begin stitching two images ...");
    // ONE pipeline thread drives the GPU (the device side of a 4096-row window takes ~13 ms); the host cores do the file
    // I/O around it in parallel slices: window k+1 is read (pread) while window k is on the device and window k-1 is
    // stored into the mapped product.  (Measured on this box's RAM disk: several GPU workers with their own contexts
    // were slower and erratic -- 2.6 s with 4, 6-12 s with 8 -- the CUDA API calls of the contexts serialise.)
    const int64_t WIN = std::max<int64_t>(64, getenv("OIP_STITCH_WINDOW") ? atoll(getenv("OIP_STITCH_WINDOW")) : 2048);
    const int64_t n_win = (lines + WIN - 1) / WIN;
    const auto t_start = std::chrono::steady_clock::now();
    const size_t in_b = (size_t)WIN * PIXELS_PER_LINE * 2, ob = (size_t)WIN * out_w * 2;
    Pinned hl0(in_b), hr0(in_b), ho0(ob), hl1(n_win > 1 ? in_b : 1), hr1(n_win > 1 ? in_b : 1), ho1(n_win > 1 ? ob : 1);
    void *hl[2] = {hl0.p, hl1.p}, *hr[2] = {hr0.p, hr1.p}, *ho[2] = {ho0.p, ho1.p};
    Fd fl(l, O_RDONLY, "cannot open file"), fr(r, O_RDONLY, "cannot open file");
    auto win_rows = [&](int64_t w) { return std::min<int64_t>(WIN, lines - w * WIN); };
    auto read_win = [&](int64_t w) {
        const size_t off = (size_t)(w * WIN) * PIXELS_PER_LINE * 2, nb = (size_t)win_rows(w) * PIXELS_PER_LINE * 2;
        parallel_io(false, fl.fd, hl[w & 1], nb, off, l);
        parallel_io(false, fr.fd, hr[w & 1], nb, off, r);
    };
    auto store_win = [&](int64_t w) { // parallel stores into the mapping
        const size_t nb = (size_t)win_rows(w) * out_w * 2;
        uint8_t *dst = out_map + data0 + (size_t)(w * WIN) * out_w * 2;
        const uint8_t *src = (const uint8_t *)ho[w & 1];
        const unsigned nt = io_threads();
        const size_t per = ((nb + nt - 1) / nt + 4095) & ~(size_t)4095;
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; ++t)
            pool.emplace_back([=]() { const size_t a0 = std::min(nb, (size_t)t * per), b0 = std::min(nb, a0 + per); if (b0 > a0) memcpy(dst + a0, src + a0, b0 - a0); });
        for (auto &th : pool) th.join();
    };
    double t_wait_rd = 0, t_gpu = 0, t_wait_wr = 0;
    auto now = []() { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a0, std::chrono::steady_clock::time_point b0) { return std::chrono::duration<double>(b0 - a0).count(); };
    std::future<void> rd = std::async(std::launch::async, read_win, (int64_t)0), wr[2];
    for (int64_t w = 0; w < n_win; ++w) {
        auto t0 = now();
        rd.get();                                                         // window w is in hl/hr[w & 1]
        if (w + 1 < n_win) rd = std::async(std::launch::async, read_win, w + 1);
        auto t1 = now();
        if (wr[w & 1].valid()) wr[w & 1].get();                           // ho[w & 1] was handed to the writers two windows ago
        auto t2 = now();
        const int64_t r0 = w * WIN, nr = win_rows(w);
        oip_pan_desc d{};
        d.n_ccd = 2; d.w = PIXELS_PER_LINE; d.total_rows = lines; d.row0 = r0; d.n_rows = nr; d.fold_half = f;
        d.section_rows = REMAP_SECTION_ROWS; d.row_guard = REMAP_ROW_GUARD;
        for (int i = 0; i < 2; ++i) {
            d.ccd[i].fmt = OIP_FMT_LE16; d.ccd[i].n_seg = 1;
            d.ccd[i].seg[0] = {i ? hr[w & 1] : hl[w & 1], r0, nr, (int64_t)PIXELS_PER_LINE * 2};
        }
        d.d_out = (uint16_t *)ho[w & 1]; d.out_pitch_px = out_w;
        oip_check(oip_pan_pipeline_host(ctx(), &d));
        auto t3 = now();
        wr[w & 1] = std::async(std::launch::async, store_win, w);
        t_wait_rd += secs(t0, t1); t_wait_wr += secs(t1, t2); t_gpu += secs(t2, t3);
    }
    for (auto &q : wr) if (q.valid()) q.get();
    const unsigned n_workers = 1;
    if (getenv("OIP_TIMING")) fprintf(stderr, "  pipeline thread: waited for reads %.3f s, for writes %.3f s, device %.3f s\n", t_wait_rd, t_wait_wr, t_gpu);
    if (getenv("OIP_TIMING"))
        fprintf(stderr, "stitch: %u workers x %lld-row windows: %.3f s for %zu B in + %zu B out\n", n_workers, (long long)WIN,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(), szl + szr, out_bytes);
    OLOG("Write stitched image to file '%s' ...", raw_out.c_str());
    OLOG("%zu bytes written.", (size_t)lines * out_w * 2);
    return 0;
}

// =============================================================================================
// downlink  (EXTENSION -- no counterpart in the reference's grammar, ref main.cpp:92-191 has auxsep / prestitch / stitch as
// three programs that talk through files, DOC/Usage.txt).  The same chain for the PAN product in one call: both CMOS
// downlinks in, stitched raster out (oip_downlink_to_stitched: AOS scan + CRC, IMTR re-framing, frame index, then the fused
// RRC + shift + stitch kernel reading the sub-images where they lie).  .IMDT / .PAN.RAW / .RRC.RAW / .PRESTT.RAW are never
// written.  The shift of CMOS-2 must be given (--dx / --dy): the estimate of `prestitch` needs the PAN rasters this command
// does not materialise.
// =============================================================================================
static int cmd_downlink(const std::vector<std::string> &av)
{
    Args a = parse(av, {{"--aos1", nullptr, false}, {"--aos2", nullptr, false}, {"--rrc1", nullptr, false}, {"--rrc2", nullptr, false},
                        {"--no-rrc", nullptr, true}, {"--dx", nullptr, false}, {"--dy", nullptr, false}, {"--fold-cols", "-c", false},
                        {"--out", "-o", false}}, 0);
    require(a, "aos1"); require(a, "aos2"); require(a, "dx"); require(a, "dy"); require(a, "fold-cols");
    const std::string f1 = a.get("aos1"), f2 = a.get("aos2");
    existing_file(f1, "--aos1"); existing_file(f2, "--aos2");
    if (a.has("rrc1")) existing_file(a.get("rrc1"), "--rrc1");
    if (a.has("rrc2")) existing_file(a.get("rrc2"), "--rrc2");
    const int fold = (int)a.geti("fold-cols", 0);
    if (fold < 2) throw parse_error(CLI_VALIDATION, "--fold-cols: fold column value too small");              // as `stitch`, ref main.cpp:166-170
    const bool do_rrc = !a.has("no-rrc");
    if (do_rrc && (!a.has("rrc1") || !a.has("rrc2"))) throw std::runtime_error("open RRC Param file failed");   // as `prestitch`
    const double dx = a.getd("dx", 0), dy = a.getd("dy", 0);
    const size_t n1 = file_size(f1), n2 = file_size(f2);
    if (n1 < 1024 || n2 < 1024) throw std::invalid_argument("AOS file too small");
    OLOG("Launching fused downlink processing (AOS separation, RRC, pre-stitch, stitch) ...");
    OLOG("    dx: %.5f, dy: %.5f (given)", dx, dy);
    DevBuf d1(n1), d2(n2);
    upload_file(f1, 0, n1, d1.p);
    upload_file(f2, 0, n2, d2.p);
    std::unique_ptr<DevBuf> kb[2];
    if (do_rrc)
        for (int i = 0; i < 2; ++i) {
            std::vector<double> v = load_rrc(a.get(i ? "rrc2" : "rrc1"), PIXELS_PER_LINE);
            kb[i].reset(new DevBuf(v.size() * 8));
            oip_check(oip_copy_h2d(ctx(), kb[i]->p, v.data(), v.size() * 8));
            oip_check(oip_ctx_sync(ctx()));                                      // v leaves scope
        }
    const oip_frame_geom g{1536, 256};                                            // ref aux_separator.h:92-93
    const int64_t frame_bytes = 49152 + 40ll * 1536 * 256 * 2 + 172, lpf = 4 * 256;
    const int out_w = oip_pan_out_width(2, PIXELS_PER_LINE, fold / 2);
    // room for every complete frame the shorter downlink can carry plus zero-filled gap frames; grown once if a stream with
    // long sequence gaps needs more (the call says how many lines it has)
    int64_t cap_rows = ((int64_t)(std::min(n1, n2) / (size_t)frame_bytes) + 8) * lpf;
    oip_downlink_stats st[2] = {};
    int64_t rows = 0;
    std::unique_ptr<DevBuf> d_out;
    for (int attempt = 0; attempt < 2; ++attempt) {
        d_out.reset();
        d_out.reset(new DevBuf((size_t)cap_rows * out_w * 2));
        oip_downlink_desc d{};
        d.n_ccd = 2; d.geom = g; d.fold_half = fold / 2; d.section_rows = REMAP_SECTION_ROWS; d.row_guard = REMAP_ROW_GUARD;
        d.ccd[0] = {(const uint8_t *)d1.p, n1, do_rrc ? (const double *)kb[0]->p : nullptr, 0, 0.0, 0.0};
        d.ccd[1] = {(const uint8_t *)d2.p, n2, do_rrc ? (const double *)kb[1]->p : nullptr, 1, dx, dy};
        d.d_out = (uint16_t *)d_out->p; d.out_pitch_px = out_w; d.out_rows_cap = cap_rows;
        const int rc = oip_downlink_to_stitched(ctx(), &d, &rows, st);
        if (rc == OIP_E_INVALID && attempt == 0) {
            const int64_t need = std::min(st[0].frames[1], st[1].frames[1]) * lpf;   // filled in before the capacity check
            if (need > cap_rows) { cap_rows = need; continue; }
        }
        oip_check(rc);
        break;
    }
    oip_check(oip_pan_check_error(ctx()));
    for (int i = 0; i < 2; ++i) {
        OLOG("CMOS-%d: %lld valid, %lld invalid, %lld empty AOS frames; %lld image transfer frames cut, %lld accepted; %lld image frames.",
             i + 1, (long long)st[i].aos[0], (long long)st[i].aos[1], (long long)st[i].aos[2], (long long)st[i].imtr[0],
             (long long)st[i].imtr[1], (long long)st[i].frames[1]);
    }
    if (rows == 0) { OLOG("No image frame data, end of job."); return 0; }
    if (rows <= REMAP_ROW_GUARD)
        OLOG("%lld lines: fewer than the %d the reference's prestitch accepts (ref imageop.h:242-244); shifted as one section.", (long long)rows,
             REMAP_ROW_GUARD + 1);
    std::string out = a.get("out");
    bool out_tiff = true;                                                                                       // as `stitch`, ref imageop.h:297-306
    if (out.empty()) out = (fs::current_path() / ("stitched_" + std::to_string(out_w) + "n16b.TIFF")).string();
    else out_tiff = lower(fs::path(out).extension().string()) == ".tiff";
    const size_t out_bytes = (size_t)rows * out_w * 2;
    size_t data0 = 0;
    if (out_tiff) data0 = (size_t)oiptiff::write_u16(out, nullptr, out_w, rows, 1, 1);   // single-band GTiff header first, ref imageop.h:316-328
    OLOG("Write stitched image to file '%s' ...", out.c_str());
    download_to_file(out, d_out->p, out_bytes, data0);
    OLOG("%zu bytes written.", out_bytes);
    return 0;
}

// =============================================================================================
// default action  (ref main.cpp:193-258, :288-317; preproc.h)
// =============================================================================================
static int cmd_default(const std::vector<std::string> &av)
{
    Args a = parse(av, {{"--pan", nullptr, false}, {"--do-rrc4pan", nullptr, true}, {"--rrc-pan", nullptr, false},
                        {"--write-rrcpan", nullptr, true}, {"--no-rrcpan", nullptr, true}, {"--mss", nullptr, false},
                        {"--no-rrc4mss", nullptr, true}, {"--rrc-msb1", nullptr, false}, {"--rrc-msb2", nullptr, false},
                        {"--rrc-msb3", nullptr, false}, {"--rrc-msb4", nullptr, false}, {"--slices", nullptr, false},
                        {"--ibc-sections", nullptr, false}, {"--ibc-threshold", nullptr, false}, {"--line-offset", nullptr, false},
                        {"--lines-section", nullptr, false}, {"--overlap-lines", nullptr, false}, {"--keep-leading", "-k", true},
                        {"--poly", nullptr, false}, {"--aligned-raw", nullptr, true}}, 0);
    if (a.has("pan")) existing_file(a.get("pan"), "--pan");
    if (a.has("mss")) existing_file(a.get("mss"), "--mss");
    if (a.has("rrc-pan") && !a.has("do-rrc4pan")) throw parse_error(CLI_REQUIRES, "--rrc-pan requires --do-rrc4pan");
    const double thr = a.getd("ibc-threshold", 0.4);
    if (thr < 0.0 || thr >= 1.0) throw parse_error(CLI_VALIDATION, "--ibc-threshold: invalid threshold value");   // ref main.cpp:233-239
    const bool rrc_mss = !a.has("no-rrc4mss");
    if (a.has("do-rrc4pan") && !a.has("rrc-pan")) throw usage_error("RRC parameter file of PAN needed");         // ref :290-292
    if (rrc_mss && !(a.has("rrc-msb1") && a.has("rrc-msb2") && a.has("rrc-msb3") && a.has("rrc-msb4")))
        throw usage_error("RRC parameter file of all MSS Bands needed");                                          // ref :293-299
    // PreProcessor::CheckFilesAttributes, ref preproc.h:552-572
    const size_t sp = file_size(a.get("pan")), sm = file_size(a.get("mss"));
    if (sp != (size_t)MSS_BANDS * sm) throw std::runtime_error("PAN file size does not match MSS file size: PAN file should be 4x as large as MSS file");
    if (sp % (PIXELS_PER_LINE * BYTES_PER_PIXEL)) throw std::runtime_error("PAN file size invalid: should be multiplies of 24576");
    double cX[8], cY[12];
    if (a.has("poly")) { // extension: skip the estimate, 4 lines "cx0 cx1 cy0 cy1 cy2"
        FILE *f = fopen(a.get("poly").c_str(), "r");
        if (!f) throw std::invalid_argument("cannot open --poly file");
        for (int b = 0; b < 4; ++b)
            if (fscanf(f, "%lf %lf %lf %lf %lf", &cX[2 * b], &cX[2 * b + 1], &cY[3 * b], &cY[3 * b + 1], &cY[3 * b + 2]) != 5) {
                fclose(f);
                throw std::invalid_argument("--poly: need 4 lines of 5 numbers");
            }
        fclose(f);
    } else {
        // PreProcessor::CalcInterBandCorrelation on the (RRC'd) PAN and MSS data, ref main.cpp:306-316, preproc.h:224-347
        const int64_t lp = (int64_t)(sp / (PIXELS_PER_LINE * BYTES_PER_PIXEL)), lm = (int64_t)(sm / (PIXELS_PER_LINE * BYTES_PER_PIXEL));
        const int wb4 = PIXELS_PER_LINE / MSS_BANDS;
        DevBuf d_pan(sp), d_ms(sm);
        upload_file(a.get("pan"), 0, sp, d_pan.p);
        upload_file(a.get("mss"), 0, sm, d_ms.p);
        if (a.has("do-rrc4pan")) {                                                                     // DoRRC4PAN, ref preproc.h:188-200
            std::vector<double> kb = load_rrc(a.get("rrc-pan"), PIXELS_PER_LINE);
            DevBuf d_kb(kb.size() * 8);
            oip_check(oip_copy_h2d(ctx(), d_kb.p, kb.data(), kb.size() * 8));
            oip_check(oip_rrc_u16(ctx(), (uint16_t *)d_pan.p, PIXELS_PER_LINE, lp, PIXELS_PER_LINE, (const double *)d_kb.p));
            oip_check(oip_ctx_sync(ctx()));
        }
        if (rrc_mss)                                                                                  // DoRRC4MSS, ref preproc.h:202-222
            for (int b = 0; b < 4; ++b) {
                char key[16]; snprintf(key, sizeof key, "rrc-msb%d", b + 1);
                existing_file(a.get(key), key);
                std::vector<double> kb = load_rrc(a.get(key), wb4);
                DevBuf d_kb(kb.size() * 8);
                oip_check(oip_copy_h2d(ctx(), d_kb.p, kb.data(), kb.size() * 8));
                oip_check(oip_rrc_u16(ctx(), (uint16_t *)d_ms.p + (size_t)b * wb4, wb4, lm, PIXELS_PER_LINE, (const double *)d_kb.p));
                oip_check(oip_ctx_sync(ctx()));
            }
        oip_ibc_config cfg{};
        cfg.slices = (int)a.geti("slices", 10); cfg.sections = (int)a.geti("ibc-sections", 5); cfg.threshold = thr;
        OLOG("Calculating inter-band correlation with %d slices in %d section(s) ...", cfg.slices, cfg.sections);
        std::vector<oip_ibc_shift> sh((size_t)4 * std::max(1, cfg.slices) * std::max(1, cfg.sections));
        oip_check(oip_inter_band_correlation(ctx(), (const uint16_t *)d_pan.p, PIXELS_PER_LINE, lp, PIXELS_PER_LINE, (const uint16_t *)d_ms.p,
                                             lm, PIXELS_PER_LINE, &cfg, sh.data(), cX, cY));
        for (int b = 0; b < 4; ++b) {
            OLOG("Doing polynomial fitting for BAND %d ...", b);
            OLOG("\tdeltaX coeff: [1] %.15f, [0] %.9f", cX[2 * b + 1], cX[2 * b]);
            OLOG("\tdeltaY coeff: [2] %.15f, [1] %.15f, [0] %.9f", cY[3 * b + 2], cY[3 * b + 1], cY[3 * b]);
        }
        OLOG("CalcInterBandCorrelation(): done.");
    }
    const int wb = PIXELS_PER_LINE / MSS_BANDS;
    const int64_t lines = (int64_t)(sm / (PIXELS_PER_LINE * BYTES_PER_PIXEL));
    oip_mss_desc m{};
    m.fmt = OIP_FMT_LE16; m.wb = wb; m.lines = lines; m.pitch_px = PIXELS_PER_LINE;
    std::vector<std::unique_ptr<DevBuf>> kbs;
    for (int b = 0; b < 4; ++b) {
        m.d_kb[b] = nullptr;
        if (rrc_mss) {
            char key[16]; snprintf(key, sizeof key, "rrc-msb%d", b + 1);
            existing_file(a.get(key), key);
            std::vector<double> kb = load_rrc(a.get(key), wb);
            kbs.emplace_back(new DevBuf(kb.size() * 8));
            oip_check(oip_copy_h2d(ctx(), kbs.back()->p, kb.data(), kb.size() * 8));
            oip_check(oip_ctx_sync(ctx()));
            m.d_kb[b] = (const double *)kbs.back()->p;
        }
    }
    memcpy(m.cX, cX, sizeof cX); memcpy(m.cY, cY, sizeof cY);
    m.lines_per_section = (int)a.geti("lines-section", 20000); m.line_offset = a.geti("line-offset", 0);
    m.overlap = (int)a.geti("overlap-lines", 520); m.keep_leading = a.has("keep-leading"); m.min_process_lines = 1500;
    const int64_t out_rows = lines - m.line_offset - (m.keep_leading ? 0 : m.overlap);
    if (out_rows <= 0) throw std::invalid_argument("Too few image lines left to process");
    DevBuf d_mss(sm), d_out((size_t)out_rows * wb * 8);
    upload_file(a.get("mss"), 0, sm, d_mss.p);
    oip_check(oip_memset_d(ctx(), d_out.p, 0, (size_t)out_rows * wb * 8));
    OLOG("Doing inter-band alignment ...");
    int64_t rows = 0;
    oip_check(oip_band_align_merge(ctx(), d_mss.p, &m, (uint16_t *)d_out.p, &rows));
    Pinned ho((size_t)out_rows * wb * 8);
    oip_check(oip_copy_d2h(ctx(), ho.p, d_out.p, (size_t)out_rows * wb * 8));
    oip_check(oip_ctx_sync(ctx()));
    if (a.has("aligned-raw")) { // extension: the CV_16UC4 memory image as it is
        const std::string out = build_output_path(a.get("mss"), ".ALIGNED", ".RAW");
        write_file(out, ho.p, (size_t)out_rows * wb * 8);
        OLOG("%lld lines aligned; written to file [%s] (CV_16UC4 layout, %d px x 4 bands per line).", (long long)rows, out.c_str(), wb);
    }
    // WriteAlignedMSS_TIFF, ref preproc.h:167-185: cv::imwrite stores a 4-channel Mat as RGBA, i.e. samples 0 and 2 swapped
    OLOG("Writing aligned MSS image as TIFF file ...");
    uint16_t *px = (uint16_t *)ho.p;
    for (size_t i = 0, n = (size_t)out_rows * wb; i < n; ++i) std::swap(px[4 * i], px[4 * i + 2]);
    const std::string tif = build_output_path(a.get("mss"), ".ALIGNED", ".TIFF");
    oiptiff::write_u16(tif, px, wb, out_rows, 4, 2, tiff_compression()); // cv::imwrite (ref preproc.h:173-177): LZW + predictor 2
    OLOG("Written to file [%s].", tif.c_str());
    return 0;
}

static void usage()
{
    puts("Optical Satellite Image Pre-Processing/Processing Utility (B200-native hot path)\n"
         "Usage: OpticalImageProcessor [OPTIONS] [SUBCOMMAND]\n\n"
         "Options: -h,--help  -v,--version  --pan --mss --rrc-msb1..4 --do-rrc4pan --rrc-pan --no-rrc4mss --line-offset\n"
         "         --lines-section --overlap-lines -k,--keep-leading --poly FILE\n"
         "Subcommands:\n"
         "  auxsep [-O,--offset N] file          Do aux & image data separation\n"
         "  prestitch --pan1 F --pan2 F [--rrc1 F --rrc2 F] [-r|--no-rrc] [-c] [-s N -l N --stitch-overlap N -e N] [--dx X --dy Y]\n"
         "  stitch --image1 F --image2 F -c,--fold-cols N -o OUT.RAW\n"
         "  downlink --aos1 F --aos2 F --rrc1 F --rrc2 F|--no-rrc --dx X --dy Y -c,--fold-cols N [-o OUT]   (extension: the three\n"
         "           steps above for the PAN product in one fused pass, no intermediate files)");
}

int main(int argc, const char *argv[])
{
    const char *lf = getenv("LOGFILE");
    g_log = fopen(lf ? lf : "oip.log", "a");
    try {
        std::vector<std::string> av(argv + 1, argv + argc);
        for (auto &t : av) {
            if (t == "-v" || t == "--version") { puts("1.1"); return 0 + 255; }       // app.exit(e) + 255, ref main.cpp:263-264
            if (t == "-h" || t == "--help") { usage(); return 0 + 255; }
        }
        try {
            int rc;
            if (!av.empty() && av[0] == "auxsep") rc = cmd_auxsep({av.begin() + 1, av.end()});
            else if (!av.empty() && av[0] == "prestitch") rc = cmd_prestitch({av.begin() + 1, av.end()});
            else if (!av.empty() && av[0] == "stitch") rc = cmd_stitch({av.begin() + 1, av.end()});
            else if (!av.empty() && av[0] == "downlink") rc = cmd_downlink({av.begin() + 1, av.end()});   // extension
            else rc = cmd_default(av);
            // every product is closed / stored by now: leave without the CUDA context teardown (0.5 - 0.9 s for nothing)
            if (g_log) fflush(g_log);
            fflush(stdout); fflush(stderr);
            _exit(rc);
        } catch (const parse_error &e) {                                               // CLI::ParseError -> app.exit(e), ref :265-266
            fprintf(stderr, "%s\nRun with --help for more information.\n", e.what());
            return e.code;
        }
    } catch (usage_error &ex) {
        printf("USAGE ERROR: %s.\n", ex.what());                                       // ref main.cpp:333-335
        return 254;
    } catch (std::exception &ex) {
        logf("E", "%s.", ex.what());                                                   // ref :336-338
        return 2;
    } catch (...) {
        logf("F", "UNKOWN FATAL ERROR OCCURED.");
        return 1;
    }
}
