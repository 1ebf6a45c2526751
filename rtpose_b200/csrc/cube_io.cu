// cube_io.cu — host side of the on-disk radar cube path: reads ONLY the ROI rows of a
// `<seq>/DZYX_npy_f16/<frame>.npy` file into a (pinned) staging slab, in place of
// `np.load(...).astype(np.float32)` + crop in CRUW_POSE_Dataset.get_cube / get_cube_phase
// (det3d/datasets/cruw_pose/cruw_pose.py:170-181, :189-192).  The x crop, the fp32 cast, the normalisation and the
// clamp stay in rtp_ingest_pack on the device, which takes the slab as a "raw" cube of extent [Z][Y][RX].
//
// File layout (numpy .npy v1/v2/v3): "\x93NUMPY" major minor, header length (u16 for v1, u32 otherwise), an ASCII
// dict {'descr': '<f2', 'fortran_order': False, 'shape': (..., RZ, RY, RX), } padded with spaces, then C-order data.
// For one leading plane l and one z the rows y0..y0+Y-1 are ONE contiguous run of Y*RX elements, so a frame is
// lead*Z preads (512 x 32 KB for the Doppler cube) instead of the whole 67 MB file.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

#include "common.cuh"

namespace {

struct Fd {
  int fd;
  explicit Fd(const char* path) : fd(::open(path, O_RDONLY | O_CLOEXEC)) {}
  ~Fd() {
    if (fd >= 0) ::close(fd);
  }
};

// full pread: loops over short reads, retries EINTR; false on error or EOF before `n` bytes
bool pread_all(int fd, char* dst, size_t n, off_t off) {
  while (n > 0) {
    ssize_t r = ::pread(fd, dst, n, off);
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    if (r == 0) {
      errno = 0;
      return false;
    }
    dst += r;
    off += r;
    n -= (size_t)r;
  }
  return true;
}

// value text of 'key' in the header dict (starts right after the colon, leading blanks skipped); npos if absent
size_t dict_value(const std::string& h, const char* key) {
  for (char q : {'\'', '"'}) {
    std::string k = std::string(1, q) + key + q;
    size_t p = h.find(k);
    if (p == std::string::npos) continue;
    p = h.find(':', p + k.size());
    if (p == std::string::npos) return p;
    ++p;
    while (p < h.size() && (h[p] == ' ' || h[p] == '\t')) ++p;
    return p;
  }
  return std::string::npos;
}

int probe_fd(int fd, const char* path, rtp_npy_info* info) {
  unsigned char pre[12];
  RTP_CHECK_ARG(pread_all(fd, (char*)pre, 10, 0), "rtp_npy: %s: shorter than a .npy preamble", path);
  RTP_CHECK_ARG(memcmp(pre, "\x93NUMPY", 6) == 0, "rtp_npy: %s: not a .npy file (bad magic)", path);
  int major = pre[6];
  RTP_CHECK_ARG(major >= 1 && major <= 3, "rtp_npy: %s: unsupported .npy version %d.%d", path, major, (int)pre[7]);
  size_t hlen, hoff;
  if (major == 1) {
    hlen = (size_t)pre[8] | ((size_t)pre[9] << 8);
    hoff = 10;
  } else {
    RTP_CHECK_ARG(pread_all(fd, (char*)pre + 10, 2, 10), "rtp_npy: %s: truncated header length", path);
    hlen = (size_t)pre[8] | ((size_t)pre[9] << 8) | ((size_t)pre[10] << 16) | ((size_t)pre[11] << 24);
    hoff = 12;
  }
  RTP_CHECK_ARG(hlen >= 2 && hlen <= (1u << 20), "rtp_npy: %s: implausible header length %zu", path, hlen);
  std::string h(hlen, '\0');
  RTP_CHECK_ARG(pread_all(fd, &h[0], hlen, (off_t)hoff), "rtp_npy: %s: truncated header", path);

  size_t p = dict_value(h, "descr");
  RTP_CHECK_ARG(p != std::string::npos && p + 1 < h.size() && (h[p] == '\'' || h[p] == '"'),
                "rtp_npy: %s: header has no simple 'descr' (structured dtypes are not cubes)", path);
  size_t e = h.find(h[p], p + 1);
  RTP_CHECK_ARG(e != std::string::npos, "rtp_npy: %s: unterminated 'descr'", path);
  std::string descr = h.substr(p + 1, e - p - 1);
  RTP_CHECK_ARG(descr.size() >= 3 && descr.size() < sizeof(info->descr), "rtp_npy: %s: odd dtype '%s'", path, descr.c_str());
  RTP_CHECK_ARG(descr[0] == '<' || descr[0] == '|' || descr[0] == '=', "rtp_npy: %s: dtype '%s' is not little-endian", path,
                descr.c_str());
  char* endp = nullptr;
  long eb = strtol(descr.c_str() + 2, &endp, 10);
  RTP_CHECK_ARG(endp && *endp == '\0' && eb >= 1 && eb <= 16, "rtp_npy: %s: odd dtype '%s'", path, descr.c_str());

  p = dict_value(h, "fortran_order");
  RTP_CHECK_ARG(p != std::string::npos, "rtp_npy: %s: header has no 'fortran_order'", path);
  int fortran = h.compare(p, 4, "True") == 0;

  p = dict_value(h, "shape");
  RTP_CHECK_ARG(p != std::string::npos && h[p] == '(', "rtp_npy: %s: header has no 'shape' tuple", path);
  e = h.find(')', p);
  RTP_CHECK_ARG(e != std::string::npos, "rtp_npy: %s: unterminated 'shape'", path);
  int nd = 0;
  int64_t count = 1;
  for (size_t i = p + 1; i < e;) {
    if (h[i] == ' ' || h[i] == ',') {
      ++i;
      continue;
    }
    RTP_CHECK_ARG(h[i] >= '0' && h[i] <= '9', "rtp_npy: %s: bad character in 'shape'", path);
    RTP_CHECK_ARG(nd < 8, "rtp_npy: %s: more than 8 dimensions", path);
    int64_t v = 0;
    while (i < e && h[i] >= '0' && h[i] <= '9') {
      v = v * 10 + (h[i] - '0');
      RTP_CHECK_ARG(v < ((int64_t)1 << 40), "rtp_npy: %s: dimension overflow", path);
      ++i;
    }
    info->shape[nd++] = v;
    count *= v;
    RTP_CHECK_ARG(count < ((int64_t)1 << 48), "rtp_npy: %s: element count overflow", path);
  }
  for (int i = nd; i < 8; ++i) info->shape[i] = 0;
  info->ndim = nd;
  info->elem_bytes = (int32_t)eb;
  info->fortran_order = fortran;
  info->data_offset = (int64_t)(hoff + hlen);
  memset(info->descr, 0, sizeof(info->descr));
  memcpy(info->descr, descr.c_str(), descr.size());
  struct stat st;
  RTP_CHECK_ARG(fstat(fd, &st) == 0, "rtp_npy: %s: fstat failed: %s", path, strerror(errno));
  info->file_bytes = (int64_t)st.st_size;
  RTP_CHECK_ARG(info->file_bytes >= info->data_offset + count * eb, "rtp_npy: %s: file holds %lld bytes, header promises %lld",
                path, (long long)info->file_bytes, (long long)(info->data_offset + count * eb));
  return 0;
}

}  // namespace

extern "C" int rtp_npy_probe(const char* path, rtp_npy_info* info) {
  RTP_CHECK_ARG(path && info, "rtp_npy_probe: null argument");
  Fd f(path);
  RTP_CHECK_ARG(f.fd >= 0, "rtp_npy_probe: cannot open %s: %s", path, strerror(errno));
  return probe_fd(f.fd, path, info);
}

extern "C" int64_t rtp_npy_roi_slab_bytes(const rtp_npy_info* info, int32_t Z, int32_t Y) {
  if (!info || info->ndim < 3 || Z <= 0 || Y <= 0) return -1;
  int64_t lead = 1;
  for (int i = 0; i < info->ndim - 3; ++i) lead *= info->shape[i];
  return lead * Z * Y * info->shape[info->ndim - 1] * info->elem_bytes;
}

extern "C" int rtp_npy_read_roi_slab(const char* path, int32_t z0, int32_t Z, int32_t y0, int32_t Y, void* dst,
                                     int64_t dst_bytes, int32_t threads) {
  RTP_CHECK_ARG(path && dst, "rtp_npy_read_roi_slab: null argument");
  Fd f(path);
  RTP_CHECK_ARG(f.fd >= 0, "rtp_npy_read_roi_slab: cannot open %s: %s", path, strerror(errno));
  rtp_npy_info in;
  int rc = probe_fd(f.fd, path, &in);
  if (rc) return rc;
  RTP_CHECK_ARG(strcmp(in.descr + 1, "f2") == 0, "rtp_npy_read_roi_slab: %s holds '%s', the cube format is float16 ('<f2')", path,
                in.descr);
  RTP_CHECK_ARG(!in.fortran_order, "rtp_npy_read_roi_slab: %s is Fortran-ordered", path);
  RTP_CHECK_ARG(in.ndim >= 3, "rtp_npy_read_roi_slab: %s has %d dimensions, a cube has at least [Z][Y][X]", path, in.ndim);
  const int64_t RZ = in.shape[in.ndim - 3], RY = in.shape[in.ndim - 2], RX = in.shape[in.ndim - 1];
  RTP_CHECK_ARG(Z > 0 && Y > 0 && z0 >= 0 && y0 >= 0 && (int64_t)z0 + Z <= RZ && (int64_t)y0 + Y <= RY,
                "rtp_npy_read_roi_slab: ROI z[%d,%d) y[%d,%d) outside the cube [%lld][%lld][%lld] of %s", z0, z0 + Z, y0, y0 + Y,
                (long long)RZ, (long long)RY, (long long)RX, path);
  const int64_t need = rtp_npy_roi_slab_bytes(&in, Z, Y);
  RTP_CHECK_ARG(dst_bytes >= need, "rtp_npy_read_roi_slab: destination holds %lld bytes, the slab needs %lld", (long long)dst_bytes,
                (long long)need);
  int64_t lead = 1;
  for (int i = 0; i < in.ndim - 3; ++i) lead *= in.shape[i];
  const int64_t run = (int64_t)Y * RX * 2;  // bytes of one (plane, z) run
  const int64_t nruns = lead * Z;
  if (nruns == 0 || run == 0) return 0;
  // a run is addressed by r = l * Z + z: file offset of row (l, z0 + z, y0), destination offset r * run
  auto file_off = [&](int64_t r) { return in.data_offset + (((r / Z) * RZ + z0 + (r % Z)) * RY + y0) * RX * 2; };

  int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(threads, 64), nruns));
  std::atomic<int64_t> next(0);
  std::atomic<int> failed(0);
  int err_no = 0;
  const int64_t grain = std::max<int64_t>(1, (1 << 20) / run);  // hand out >= 1 MB of runs at a time
  auto work = [&]() {
    for (;;) {
      int64_t r0 = next.fetch_add(grain);
      if (r0 >= nruns || failed.load()) return;
      int64_t r1 = std::min(nruns, r0 + grain);
      for (int64_t r = r0; r < r1; ++r) {
        // when the y range spans whole planes consecutive z runs are adjacent in the file: merge them into one pread
        int64_t m = 1;
        if (Y == RY)
          while (r + m < r1 && (r + m) % Z != 0) ++m;
        if (!pread_all(f.fd, (char*)dst + r * run, (size_t)(run * m), (off_t)file_off(r))) {
          if (!failed.exchange(1)) err_no = errno;
          return;
        }
        r += m - 1;
      }
    }
  };
  if (nt == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    pool.reserve(nt - 1);
    try {
      for (int i = 1; i < nt; ++i) pool.emplace_back(work);
    } catch (...) {
      // the process is out of threads: whatever helpers exist plus this thread drain the queue (no exception may cross the C ABI)
    }
    work();
    for (auto& t : pool) t.join();
  }
  if (failed.load()) {
    rtp_set_error("rtp_npy_read_roi_slab: read of %s failed: %s", path, err_no ? strerror(err_no) : "unexpected end of file");
    return -2;
  }
  return 0;
}
