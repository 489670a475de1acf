// lib_core.inl -- errors, device info, npy/HugeConst loader (host side).
// HugeConst follows UO/huge_const_op/huge_const_op.cc:85-226 (checks, error codes, one cached
// device copy); the npy header grammar is NumPy's format 1.0/2.0/3.0.

namespace nann {

std::atomic<uint64_t> g_launches{0};
static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}
nann_status fail(nann_status code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

nann_status require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(NANN_FAILED_PRECONDITION,
                "libnann_b200 needs a CUDA device (sm_100a); there is no CPU fallback (%s)",
                e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
  }
  // the library carries sm_100a code only: on any other GPU every launch would fail with "no kernel image"
  static std::atomic<int> cc_ok[64];   // per device: 0 unknown, 1 ok, 2 wrong architecture
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return fail(NANN_FAILED_PRECONDITION, "no current CUDA device"); }
  int state = cc_ok[dev & 63].load(std::memory_order_relaxed);
  if (state == 0) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { cudaGetLastError(); major = 0; }
    state = major == 10 ? 1 : 2;
    cc_ok[dev & 63].store(state, std::memory_order_relaxed);
  }
  if (state != 1)
    return fail(NANN_FAILED_PRECONDITION, "libnann_b200 is built for sm_100a (B200) only; device %d has another architecture", dev);
  return NANN_OK;
}

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

static const int kDtypeSize[5] = {2, 4, 8, 4, 8};
static const char* kDtypeDescr[5] = {"<f2", "<f4", "<f8", "<i4", "<i8"};

struct NpyHeader {
  int dtype = -1;
  bool fortran = false;
  std::vector<int64_t> shape;
  size_t data_offset = 0;
  std::string descr;
};

// returns OK / NOT_FOUND / INTERNAL(malformed)
static nann_status npy_read_header(const char* path, NpyHeader* h, FILE** keep_open) {
  FILE* f = fopen(path, "rb");
  if (!f) return fail(NANN_NOT_FOUND, "Fail to open file: %s", path);  // huge_const_op.cc:94-96
  unsigned char m[12];
  auto bad = [&](const char* why) {
    fclose(f);
    return fail(NANN_INTERNAL, "npy header of %s: %s", path, why);
  };
  if (fread(m, 1, 8, f) != 8 || memcmp(m, "\x93NUMPY", 6) != 0) return bad("bad magic");
  size_t hlen;
  if (m[6] == 1) {
    if (fread(m + 8, 1, 2, f) != 2) return bad("truncated");
    hlen = m[8] | ((size_t)m[9] << 8);
    h->data_offset = 10 + hlen;
  } else if (m[6] == 2 || m[6] == 3) {
    if (fread(m + 8, 1, 4, f) != 4) return bad("truncated");
    hlen = m[8] | ((size_t)m[9] << 8) | ((size_t)m[10] << 16) | ((size_t)m[11] << 24);
    h->data_offset = 12 + hlen;
  } else {
    return bad("unsupported npy version");
  }
  std::string hdr(hlen, '\0');
  if (fread(&hdr[0], 1, hlen, f) != hlen) return bad("truncated header");
  auto find_val = [&](const char* key) -> size_t {
    size_t p = hdr.find(key);
    if (p == std::string::npos) return p;
    p = hdr.find(':', p);
    if (p == std::string::npos) return p;
    ++p;
    while (p < hdr.size() && hdr[p] == ' ') ++p;
    return p;
  };
  size_t p = find_val("'descr'");
  if (p == std::string::npos || (hdr[p] != '\'' && hdr[p] != '"')) return bad("no descr");
  size_t e = hdr.find(hdr[p], p + 1);
  h->descr = hdr.substr(p + 1, e - p - 1);
  h->dtype = -1;
  for (int i = 0; i < 5; ++i)
    if (h->descr == kDtypeDescr[i] || (h->descr.size() == 3 && h->descr[0] == '=' &&
                                       h->descr.compare(1, 2, kDtypeDescr[i] + 1) == 0))
      h->dtype = i;
  p = find_val("'fortran_order'");
  if (p == std::string::npos) return bad("no fortran_order");
  h->fortran = hdr.compare(p, 4, "True") == 0;
  p = find_val("'shape'");
  if (p == std::string::npos || hdr[p] != '(') return bad("no shape");
  ++p;
  h->shape.clear();
  while (p < hdr.size() && hdr[p] != ')') {
    while (p < hdr.size() && (hdr[p] == ' ' || hdr[p] == ',')) ++p;
    if (hdr[p] == ')') break;
    char* endp = nullptr;
    long long v = strtoll(hdr.c_str() + p, &endp, 10);
    if (endp == hdr.c_str() + p) return bad("bad shape");
    h->shape.push_back(v);
    p = endp - hdr.c_str();
  }
  if (keep_open) *keep_open = f; else fclose(f);
  return NANN_OK;
}

}  // namespace nann

using namespace nann;

struct nann_huge_const {
  int dtype = 0;
  std::vector<int64_t> shape;
  int64_t bytes = 0;
  void* host = nullptr;    // pinned when a device is present, malloc otherwise
  bool pinned = false;
  void* dev = nullptr;
  int device = -1;
};

extern "C" {

int nann_abi_version(void) { return NANN_B200_ABI_VERSION; }
const char* nann_last_error(void) { return g_err.c_str(); }
uint64_t nann_kernel_launch_count(void) { return g_launches.load(); }

nann_status nann_device_info(int device, int* device_count, int* sm_count, int64_t* hbm_bytes,
                             int* cc_major, int* cc_minor) {
  NANN_TRY(require_device());
  int n = 0;
  NANN_CUDA(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return fail(NANN_INVALID_ARGUMENT, "device %d of %d", device, n);
  cudaDeviceProp pr;
  NANN_CUDA(cudaGetDeviceProperties(&pr, device));
  if (device_count) *device_count = n;
  if (sm_count) *sm_count = pr.multiProcessorCount;
  if (hbm_bytes) *hbm_bytes = (int64_t)pr.totalGlobalMem;
  if (cc_major) *cc_major = pr.major;
  if (cc_minor) *cc_minor = pr.minor;
  return NANN_OK;
}

nann_status nann_npy_peek(const char* path, int* dtype, int* rank, int64_t* shape8) {
  NpyHeader h;
  NANN_TRY(npy_read_header(path, &h, nullptr));
  if (h.shape.size() > 8) return fail(NANN_UNIMPLEMENTED, "rank %zu > 8", h.shape.size());
  if (dtype) *dtype = h.dtype;
  if (rank) *rank = (int)h.shape.size();
  if (shape8) for (size_t i = 0; i < h.shape.size(); ++i) shape8[i] = h.shape[i];
  return NANN_OK;
}

nann_status nann_huge_const_create(const char* path, int dtype, const int64_t* shape, int rank,
                                   int device, nann_huge_const_t** out) {
  if (!out || !path) return fail(NANN_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (dtype < 0 || dtype > 4) return fail(NANN_UNIMPLEMENTED, "Unsupported DataType.");  // :143-146
  NpyHeader h;
  FILE* f = nullptr;
  NANN_TRY(npy_read_header(path, &h, &f));
  auto bail = [&](nann_status s) { fclose(f); return s; };
  if (h.fortran) return bail(fail(NANN_UNIMPLEMENTED, "Fortran order NOT supported."));  // :105-107
  for (size_t i = 0; i < h.shape.size(); ++i)                                             // :110-115
    if ((int)i >= rank || h.shape[i] != shape[i])
      return bail(fail(NANN_INTERNAL, "attr_shape and np_shape NOT match in dim %zu", i));
  if (h.dtype != dtype)                                                                   // :118-147
    return bail(fail(NANN_INTERNAL, "DataType mismatch: %s!=%s", h.descr.c_str(), kDtypeDescr[dtype]));
  int64_t count = 1;
  for (int i = 0; i < rank; ++i) count *= shape[i];
  auto* hc = new nann_huge_const();
  hc->dtype = dtype;
  hc->shape.assign(shape, shape + rank);
  hc->bytes = count * kDtypeSize[dtype];
  hc->device = device;
  const size_t alloc_bytes = (size_t)(hc->bytes > 0 ? hc->bytes : 1);
  if (device >= 0) {
    if (require_device() != NANN_OK) { delete hc; return bail(NANN_FAILED_PRECONDITION); }
    cudaSetDevice(device);
    if (cudaMallocHost(&hc->host, alloc_bytes) == cudaSuccess) hc->pinned = true;
    else { cudaGetLastError(); hc->host = nullptr; }
  }
  if (!hc->host) hc->host = malloc(alloc_bytes);
  if (!hc->host) { delete hc; return bail(fail(NANN_RESOURCE_EXHAUSTED, "OOM reading %s", path)); }
  fseek(f, (long)h.data_offset, SEEK_SET);
  size_t got = fread(hc->host, 1, (size_t)hc->bytes, f);                                  // :181
  fclose(f);
  if ((int64_t)got != hc->bytes) {
    nann_huge_const_destroy(hc);
    return fail(NANN_INTERNAL, "%s: payload shorter than shape", path);
  }
  if (device >= 0) {  // one blocking H2D, cached for the kernel's lifetime (:184-226)
    cudaError_t e = cudaMalloc(&hc->dev, alloc_bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      nann_huge_const_destroy(hc);
      return fail(NANN_RESOURCE_EXHAUSTED, "OOM when allocating tensor of %lld bytes", (long long)alloc_bytes);
    }
    e = cudaMemcpy(hc->dev, hc->host, (size_t)hc->bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      nann_huge_const_destroy(hc);
      return fail(NANN_INTERNAL, "H2D copy failed: %s", cudaGetErrorString(e));
    }
  }
  *out = hc;
  return NANN_OK;
}
const void* nann_huge_const_host(const nann_huge_const_t* h) { return h ? h->host : nullptr; }
const void* nann_huge_const_device(const nann_huge_const_t* h) { return h ? h->dev : nullptr; }
int64_t nann_huge_const_bytes(const nann_huge_const_t* h) { return h ? h->bytes : 0; }
void nann_huge_const_destroy(nann_huge_const_t* h) {
  if (!h) return;
  if (h->dev) cudaFree(h->dev);
  if (h->host) { if (h->pinned) cudaFreeHost(h->host); else free(h->host); }
  delete h;
}

}  // extern "C"
