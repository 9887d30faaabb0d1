// ORACLE — TEST INFRASTRUCTURE ONLY.  Runs code of the REFERENCE'S SHIPPED EXECUTABLE (/root/reference/bin/arch_x64/sift).
//
// That file is a non-PIE, unstripped x86-64 ELF built by the reference's author with GCC 7 against the real Vigra 1.11:
// every Vigra template the path uses (Kernel1D::initGaussian, the reflect line convolution, resizeImageNoInterpolation,
// linalg::inverse / linearSolve / qrDecomposition) and the reference's own sift::alg:: / sift::Sift:: functions are compiled
// into it.  It cannot be started here (it wants Vigra-impex, OpenCV, Boost as shared libraries), but none of the functions on
// the path calls into those.  So this helper maps the executable's two LOAD segments at their link addresses, binds its
// PLT/GOT to this process's libc / libm / libstdc++ (anything else becomes a stub that names the symbol and exits), registers
// its unwind tables, looks the functions up in its symbol table and CALLS THEM with argument objects laid out as Vigra and the
// reference's headers lay them out.  What comes back is what the reference's own build computes, Vigra included — the part
// that oracle/_ref (reference sources over stand-in headers) cannot pin.
//
// usage: refbin_run <executable> stages <in> <out>
//   in : i32 w, h, dpe, octaves; f32 sigma, k; i32 subpixel; w*h f32 (row-major)
//   out: the Gaussian and DoG pyramids of Sift::_createDOGs, the candidate list after _findScaleSpaceExtrema +
//        _eliminateEdgeResponses (flags included), and the result vector of Sift::calculate on a fresh object
//        (see dump_* below for the record formats)
//        refbin_run <executable> calculate <in> <out> same input; only Sift::calculate, out: the result vector (for frames on which
//                                                     the stage-by-stage run, which does the work twice, would take too long)
//        refbin_run <executable> time <in> <out>      same input as `stages`; only Sift::calculate, out: f64 seconds, i32 keypoints
//        refbin_run <executable> unit <in> <out>      alg::convolveWithGauss / reduceToNextLevel / increaseToNextLevel on one image
//        refbin_run <executable> eliminate <in> <out> Sift::_eliminateEdgeResponses (Vigra's inverse + linearSolve inside) on three
//                                                     caller-supplied DoG layers and a list of points; out: one flag per point
//        refbin_run <executable> vertex <in> <out>    alg::vertexParabola on n point triples (the rank-deficient least squares)
//        refbin_run <executable> peaks <in> <out>     Sift::_findPeaks on n 36-bin histograms; out per histogram: i32 count, the set
//
// Never part of the product; tests/test_refbin_pin.py is its only user and runs where /root/reference exists.
#include <dlfcn.h>
#include <elf.h>
#include <sys/mman.h>
#include <unistd.h>

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <array>
#include <chrono>
#include <new>
#include <set>
#include <string>
#include <vector>

// the reference's own result type (no Vigra inside): layout by construction
#include "interestpoint.hpp"

extern "C" void __register_frame(void*);

namespace {

[[noreturn]] void die(const char* what, const char* arg = "") {
    std::fprintf(stderr, "refbin_run: %s %s\n", what, arg);
    _exit(70);
}

extern "C" [[noreturn]] void refbin_unresolved(const char* name) {
    std::fprintf(stderr, "refbin_run: the executable called %s, which is not available here\n", name);
    _exit(77);
}

struct Exe {
    std::vector<unsigned char> file;
    std::map<std::string, uint64_t> sym;
    unsigned char* thunks = nullptr;
    size_t thunk_used = 0;

    void* trap_for(const char* name) {
        if (!thunks) {
            thunks = (unsigned char*)mmap(nullptr, 1 << 16, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (thunks == MAP_FAILED) die("mmap thunks");
        }
        if (thunk_used + 32 > (1 << 16)) die("too many unresolved symbols");
        unsigned char* t = thunks + thunk_used;
        thunk_used += 32;
        const uint64_t n = (uint64_t)strdup(name), h = (uint64_t)&refbin_unresolved;
        t[0] = 0x48; t[1] = 0xBF; std::memcpy(t + 2, &n, 8);      // mov rdi, name
        t[10] = 0x48; t[11] = 0xB8; std::memcpy(t + 12, &h, 8);   // mov rax, handler
        t[20] = 0xFF; t[21] = 0xE0;                               // jmp rax
        return t;
    }

    void load(const char* path) {
        FILE* f = std::fopen(path, "rb");
        if (!f) die("cannot open", path);
        std::fseek(f, 0, SEEK_END);
        file.resize((size_t)std::ftell(f));
        std::fseek(f, 0, SEEK_SET);
        if (std::fread(file.data(), 1, file.size(), f) != file.size()) die("short read", path);
        std::fclose(f);
        const Elf64_Ehdr* eh = (const Elf64_Ehdr*)file.data();
        if (std::memcmp(eh->e_ident, ELFMAG, SELFMAG) != 0 || eh->e_ident[EI_CLASS] != ELFCLASS64 || eh->e_machine != EM_X86_64 || eh->e_type != ET_EXEC)
            die("not a non-PIE x86-64 executable:", path);
        const Elf64_Phdr* ph = (const Elf64_Phdr*)(file.data() + eh->e_phoff);
        const Elf64_Dyn* dyn = nullptr;
        const long page = sysconf(_SC_PAGESIZE);
        for (int i = 0; i < eh->e_phnum; ++i) {
            if (ph[i].p_type == PT_TLS) die("the executable has thread-local storage: not handled");
            if (ph[i].p_type == PT_DYNAMIC) dyn = (const Elf64_Dyn*)ph[i].p_vaddr;
            if (ph[i].p_type != PT_LOAD) continue;
            const uint64_t lo = ph[i].p_vaddr & ~(uint64_t)(page - 1);
            const uint64_t hi = (ph[i].p_vaddr + ph[i].p_memsz + page - 1) & ~(uint64_t)(page - 1);
            void* got = mmap((void*)lo, hi - lo, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS | MAP_FIXED_NOREPLACE, -1, 0);
            if (got != (void*)lo) die("cannot map a segment at its link address (is this helper built as PIE?)");
            std::memcpy((void*)ph[i].p_vaddr, file.data() + ph[i].p_offset, ph[i].p_filesz);
        }
        if (!dyn) die("no dynamic section");
        const Elf64_Sym* dynsym = nullptr;
        const char* dynstr = nullptr;
        const Elf64_Rela *rela = nullptr, *jmprel = nullptr;
        size_t relasz = 0, pltsz = 0;
        for (const Elf64_Dyn* d = dyn; d->d_tag != DT_NULL; ++d) {
            if (d->d_tag == DT_SYMTAB) dynsym = (const Elf64_Sym*)d->d_un.d_ptr;
            if (d->d_tag == DT_STRTAB) dynstr = (const char*)d->d_un.d_ptr;
            if (d->d_tag == DT_RELA) rela = (const Elf64_Rela*)d->d_un.d_ptr;
            if (d->d_tag == DT_RELASZ) relasz = d->d_un.d_val;
            if (d->d_tag == DT_JMPREL) jmprel = (const Elf64_Rela*)d->d_un.d_ptr;
            if (d->d_tag == DT_PLTRELSZ) pltsz = d->d_un.d_val;
        }
        if (!dynsym || !dynstr) die("no dynamic symbols");
        auto apply = [&](const Elf64_Rela* r, size_t bytes) {
            for (size_t i = 0; r && i < bytes / sizeof(Elf64_Rela); ++i) {
                const Elf64_Sym& s = dynsym[ELF64_R_SYM(r[i].r_info)];
                const char* name = dynstr + s.st_name;
                void* here = *name ? dlsym(RTLD_DEFAULT, name) : nullptr;
                switch (ELF64_R_TYPE(r[i].r_info)) {
                    case R_X86_64_JUMP_SLOT:
                        *(uint64_t*)r[i].r_offset = (uint64_t)(here ? here : trap_for(name));
                        break;
                    case R_X86_64_GLOB_DAT:
                        *(uint64_t*)r[i].r_offset = (uint64_t)here;   // weak undefined (__gmon_start__ ...) stay null
                        break;
                    case R_X86_64_COPY:                               // vtables / typeinfo of libstdc++ classes, std::cout ...
                        if (here) std::memcpy((void*)r[i].r_offset, here, s.st_size);
                        break;
                    case R_X86_64_64:
                        *(uint64_t*)r[i].r_offset = (uint64_t)here + (uint64_t)r[i].r_addend;
                        break;
                    default:
                        die("relocation type not handled");
                }
            }
        };
        apply(rela, relasz);
        apply(jmprel, pltsz);
        // the full symbol table and the unwind tables come from the section headers
        const Elf64_Shdr* sh = (const Elf64_Shdr*)(file.data() + eh->e_shoff);
        const char* shstr = (const char*)file.data() + sh[eh->e_shstrndx].sh_offset;
        for (int i = 0; i < eh->e_shnum; ++i) {
            if (sh[i].sh_type == SHT_SYMTAB) {
                const Elf64_Sym* st = (const Elf64_Sym*)(file.data() + sh[i].sh_offset);
                const char* str = (const char*)file.data() + sh[sh[i].sh_link].sh_offset;
                for (size_t k = 0; k < sh[i].sh_size / sizeof(Elf64_Sym); ++k)
                    if (ELF64_ST_TYPE(st[k].st_info) == STT_FUNC && st[k].st_value) sym[str + st[k].st_name] = st[k].st_value;
            }
            if (std::strcmp(shstr + sh[i].sh_name, ".eh_frame") == 0) __register_frame((void*)sh[i].sh_addr);
        }
        if (sym.empty()) die("the executable is stripped");
    }

    template <typename F>
    F fn(const char* mangled) const {
        auto it = sym.find(mangled);
        if (it == sym.end()) die("symbol not found:", mangled);
        return reinterpret_cast<F>(it->second);
    }
};

// ---- the argument objects, laid out as the executable's headers lay them out -----------------------------------------
// vigra::MultiArray<2, float> (Vigra 1.11 multi_array.hxx): MultiArrayView { TinyVector<ptrdiff_t,2> m_shape, m_stride; float* m_ptr }
// followed by the (empty) allocator member; 48 bytes.  The user-provided destructor makes the type non-trivial for the purposes
// of calls, so it travels by invisible reference exactly like the real class.
struct VArr {
    long shape[2];
    long stride[2];
    float* ptr;
    char alloc;
    VArr() : shape{0, 0}, stride{0, 0}, ptr(nullptr), alloc(0) {}
    VArr(const float* src, int w, int h) : shape{w, h}, stride{1, w}, alloc(0) {
        ptr = static_cast<float*>(::operator new(sizeof(float) * (size_t)w * (size_t)h));   // std::allocator<float>: the callee may free it
        std::memcpy(ptr, src, sizeof(float) * (size_t)w * (size_t)h);
    }
    VArr(const VArr&) = delete;
    ~VArr() {}   // leaked on purpose: the process is short-lived
    float at(long x, long y) const { return ptr[x * stride[0] + y * stride[1]]; }
};
static_assert(sizeof(VArr) == 48, "vigra::MultiArray<2, float> is 48 bytes");

// sift::OctaveElem (octaveelem.hpp): f32 scale; MultiArray img
struct ROctaveElem {
    float scale;
    VArr img;
};
static_assert(sizeof(ROctaveElem) == 56, "sift::OctaveElem is 56 bytes");

// sift::Matrix<T> (matrix.hpp:21-27): u16 _width, _height; std::shared_ptr<T> _data; element (x, y) at x * _height + y
struct RMatrix {
    uint16_t width = 0, height = 0;
    std::shared_ptr<ROctaveElem> data;
    const ROctaveElem& at(int x, int y) const { return data.get()[x * height + y]; }
};
static_assert(sizeof(RMatrix) == 24, "sift::Matrix<T> is 24 bytes");

// sift::Sift (sift.hpp:17-58): subpixel, _sigma, _k, _dogsPerEpoch, _octaves, _gaussians, _magnitudes, _orientations.
// Its constructor is inline (sift.hpp:66-71) and only stores the five scalars; the matrices start default-constructed.
struct RSift {
    bool subpixel;
    float sigma, k;
    uint16_t dpe, octaves;
    RMatrix gaussians, magnitudes, orientations;
};
static_assert(sizeof(RSift) == 88 && offsetof(RSift, gaussians) == 16, "sift::Sift layout");

typedef std::vector<sift::InterestPoint> Points;

void put(FILE* f, const void* p, size_t n) {
    if (std::fwrite(p, 1, n, f) != n) die("write failed");
}
template <typename T>
void put(FILE* f, T v) { put(f, &v, sizeof v); }

void dump_image(FILE* f, const VArr& a, float scale) {
    put<int32_t>(f, (int32_t)a.shape[0]);
    put<int32_t>(f, (int32_t)a.shape[1]);
    put<float>(f, scale);
    for (long y = 0; y < a.shape[1]; ++y)
        for (long x = 0; x < a.shape[0]; ++x) put<float>(f, a.at(x, y));
}

void dump_points(FILE* f, const Points& v, bool with_desc) {
    put<int32_t>(f, (int32_t)v.size());
    for (const sift::InterestPoint& p : v) {
        put<uint16_t>(f, p.loc.x); put<uint16_t>(f, p.loc.y); put<uint16_t>(f, p.octave); put<uint16_t>(f, p.index);
        put<float>(f, p.scale);
        put<float>(f, with_desc ? p.orientation : 0.0f);
        put<uint8_t>(f, p.filtered ? 1 : 0);
        if (with_desc) {
            put<int32_t>(f, (int32_t)p.descriptors.size());
            put(f, p.descriptors.data(), sizeof(float) * p.descriptors.size());
        }
    }
}

RSift* make_sift(int dpe, int octaves, float sigma, float k, bool subpixel) {
    RSift* s = new RSift();   // leaked: its shared_ptrs would release through the executable's control blocks, which is fine, just needless
    s->subpixel = subpixel; s->sigma = sigma; s->k = k; s->dpe = (uint16_t)dpe; s->octaves = (uint16_t)octaves;
    return s;
}

std::vector<unsigned char> slurp(const char* path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) die("cannot open", path);
    std::vector<unsigned char> b;
    unsigned char tmp[1 << 16];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) b.insert(b.end(), tmp, tmp + n);
    std::fclose(f);
    return b;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc != 5) die("usage: refbin_run <executable> stages|unit <in> <out>");
    Exe exe;
    exe.load(argv[1]);
    const std::string cmd = argv[2];
    const std::vector<unsigned char> in = slurp(argv[3]);
    FILE* out = std::fopen(argv[4], "wb");
    if (!out) die("cannot write", argv[4]);

    typedef VArr (*ImgFn)(const VArr&, float);
    typedef RMatrix (*DogsFn)(RSift*, VArr&);
    typedef void (*ExtremaFn)(const RSift*, const RMatrix&, Points&);
    typedef void (*ElimFn)(const RSift*, Points&, const RMatrix&);
    typedef Points (*CalcFn)(RSift*, VArr&);

    try {
        if (cmd == "stages") {
            struct Hdr { int32_t w, h, dpe, octaves; float sigma, k; int32_t subpixel; };
            if (in.size() < sizeof(Hdr)) die("short input");
            Hdr hd;
            std::memcpy(&hd, in.data(), sizeof hd);
            if (in.size() != sizeof(Hdr) + sizeof(float) * (size_t)hd.w * (size_t)hd.h) die("input size does not match its header");
            const float* px = reinterpret_cast<const float*>(in.data() + sizeof(Hdr));
            // 1. the pyramid and the candidate list, stage by stage (what Sift::calculate does first, sift.cpp:19-35)
            RSift* s = make_sift(hd.dpe, hd.octaves, hd.sigma, hd.k, hd.subpixel != 0);
            VArr* img = new VArr(px, hd.w, hd.h);
            if (hd.subpixel) {
                img = new VArr(exe.fn<ImgFn>("_ZN4sift3alg19increaseToNextLevelERKN5vigra10MultiArrayILj2EfSaIfEEEf")(*img, 1.0f));  // sift.cpp:20-21
            }
            RMatrix* dogs = new RMatrix(exe.fn<DogsFn>("_ZN4sift4Sift11_createDOGsERN5vigra10MultiArrayILj2EfSaIfEEE")(s, *img));
            put<int32_t>(out, hd.octaves);
            put<int32_t>(out, hd.dpe);
            for (int o = 0; o < hd.octaves; ++o) {
                for (int i = 0; i < hd.dpe + 1; ++i) dump_image(out, s->gaussians.at(o, i).img, s->gaussians.at(o, i).scale);
                for (int i = 0; i < hd.dpe; ++i) dump_image(out, dogs->at(o, i).img, dogs->at(o, i).scale);
            }
            Points* cands = new Points();
            exe.fn<ExtremaFn>("_ZNK4sift4Sift22_findScaleSpaceExtremaERKNS_6MatrixINS_10OctaveElemEEERSt6vectorINS_13InterestPointESaIS7_EE")(s, *dogs, *cands);
            exe.fn<ElimFn>("_ZNK4sift4Sift23_eliminateEdgeResponsesERSt6vectorINS_13InterestPointESaIS2_EERKNS_6MatrixINS_10OctaveElemEEE")(s, *cands, *dogs);
            dump_points(out, *cands, false);
            // 2. the whole of Sift::calculate on a fresh object and a fresh copy of the input
            RSift* s2 = make_sift(hd.dpe, hd.octaves, hd.sigma, hd.k, hd.subpixel != 0);
            VArr* img2 = new VArr(px, hd.w, hd.h);
            Points* res = new Points(exe.fn<CalcFn>("_ZN4sift4Sift9calculateERN5vigra10MultiArrayILj2EfSaIfEEE")(s2, *img2));
            dump_points(out, *res, true);
        } else if (cmd == "calculate") {
            struct Hdr { int32_t w, h, dpe, octaves; float sigma, k; int32_t subpixel; };
            Hdr hd;
            if (in.size() < sizeof(Hdr)) die("short input");
            std::memcpy(&hd, in.data(), sizeof hd);
            if (in.size() != sizeof(Hdr) + sizeof(float) * (size_t)hd.w * (size_t)hd.h) die("input size does not match its header");
            RSift* s = make_sift(hd.dpe, hd.octaves, hd.sigma, hd.k, hd.subpixel != 0);
            VArr* img = new VArr(reinterpret_cast<const float*>(in.data() + sizeof(Hdr)), hd.w, hd.h);
            Points* res = new Points(exe.fn<CalcFn>("_ZN4sift4Sift9calculateERN5vigra10MultiArrayILj2EfSaIfEEE")(s, *img));
            dump_points(out, *res, true);
        } else if (cmd == "time") {
            struct Hdr { int32_t w, h, dpe, octaves; float sigma, k; int32_t subpixel; };
            Hdr hd;
            if (in.size() < sizeof(Hdr)) die("short input");
            std::memcpy(&hd, in.data(), sizeof hd);
            if (in.size() != sizeof(Hdr) + sizeof(float) * (size_t)hd.w * (size_t)hd.h) die("input size does not match its header");
            RSift* s = make_sift(hd.dpe, hd.octaves, hd.sigma, hd.k, hd.subpixel != 0);
            VArr* img = new VArr(reinterpret_cast<const float*>(in.data() + sizeof(Hdr)), hd.w, hd.h);
            const auto t0 = std::chrono::steady_clock::now();
            Points* res = new Points(exe.fn<CalcFn>("_ZN4sift4Sift9calculateERN5vigra10MultiArrayILj2EfSaIfEEE")(s, *img));
            put<double>(out, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            put<int32_t>(out, (int32_t)res->size());
        } else if (cmd == "unit") {
            struct Hdr { int32_t w, h; float sigma; };
            Hdr hd;
            if (in.size() < sizeof(Hdr)) die("short input");
            std::memcpy(&hd, in.data(), sizeof hd);
            if (in.size() != sizeof(Hdr) + sizeof(float) * (size_t)hd.w * (size_t)hd.h) die("input size does not match its header");
            const float* px = reinterpret_cast<const float*>(in.data() + sizeof(Hdr));
            VArr* img = new VArr(px, hd.w, hd.h);
            const char* names[3] = {"_ZN4sift3alg17convolveWithGaussERKN5vigra10MultiArrayILj2EfSaIfEEEf",
                                    "_ZN4sift3alg17reduceToNextLevelERKN5vigra10MultiArrayILj2EfSaIfEEEf",
                                    "_ZN4sift3alg19increaseToNextLevelERKN5vigra10MultiArrayILj2EfSaIfEEEf"};
            for (const char* n : names) {
                VArr* r = new VArr(exe.fn<ImgFn>(n)(*img, hd.sigma));
                dump_image(out, *r, hd.sigma);
            }
        } else if (cmd == "eliminate") {
            // in: i32 w, h, n; three w*h f32 layers (row-major); n x (u16 x, u16 y)
            struct Hdr { int32_t w, h, n; };
            Hdr hd;
            if (in.size() < sizeof(Hdr)) die("short input");
            std::memcpy(&hd, in.data(), sizeof hd);
            const size_t px = (size_t)hd.w * (size_t)hd.h;
            if (in.size() != sizeof(Hdr) + 3 * sizeof(float) * px + 4 * (size_t)hd.n) die("input size does not match its header");
            const float* layers = reinterpret_cast<const float*>(in.data() + sizeof(Hdr));
            const uint16_t* xy = reinterpret_cast<const uint16_t*>(in.data() + sizeof(Hdr) + 3 * sizeof(float) * px);
            // sift::Matrix<OctaveElem>(1, 3): one octave, three DoGs, element (0, i) at index i
            ROctaveElem* elems = static_cast<ROctaveElem*>(::operator new(sizeof(ROctaveElem) * 3));
            for (int i = 0; i < 3; ++i) {
                elems[i].scale = 1.0f + (float)i;
                new (&elems[i].img) VArr(layers + (size_t)i * px, hd.w, hd.h);
            }
            RMatrix* dogs = new RMatrix();
            dogs->width = 1; dogs->height = 3;
            dogs->data = std::shared_ptr<ROctaveElem>(elems, [](ROctaveElem*) {});
            Points* pts = new Points();
            for (int j = 0; j < hd.n; ++j) pts->push_back(sift::InterestPoint(sift::Point<u16_t, u16_t>(xy[2 * j], xy[2 * j + 1]), 2.0f, 0, 1));
            RSift* s = make_sift(3, 1, 1.6f, 1.4142135f, false);
            exe.fn<ElimFn>("_ZNK4sift4Sift23_eliminateEdgeResponsesERSt6vectorINS_13InterestPointESaIS2_EERKNS_6MatrixINS_10OctaveElemEEE")(s, *pts, *dogs);
            for (const sift::InterestPoint& p : *pts) put<uint8_t>(out, p.filtered ? 1 : 0);
        } else if (cmd == "vertex") {
            // in: i32 n; n x (u16 lx, f32 ly, u16 px, f32 py, u16 rx, f32 ry) packed as 6 f32 (x values as floats holding integers)
            typedef float (*VertexFn)(const sift::Point<u16_t, f32_t>&, const sift::Point<u16_t, f32_t>&, const sift::Point<u16_t, f32_t>&);
            int32_t n;
            std::memcpy(&n, in.data(), 4);
            if (in.size() != 4 + (size_t)n * 24) die("input size does not match its header");
            const float* v = reinterpret_cast<const float*>(in.data() + 4);
            VertexFn f = exe.fn<VertexFn>("_ZN4sift3alg14vertexParabolaERKNS_5PointItfEES4_S4_");
            for (int j = 0; j < n; ++j) {
                const sift::Point<u16_t, f32_t> a((u16_t)v[6 * j], v[6 * j + 1]), b((u16_t)v[6 * j + 2], v[6 * j + 3]), c((u16_t)v[6 * j + 4], v[6 * j + 5]);
                put<float>(out, f(a, b, c));
            }
        } else if (cmd == "peaks") {
            // in: i32 n; n x 36 f32
            typedef std::set<float> (*PeaksFn)(const RSift*, const std::array<float, 36>&);
            int32_t n;
            std::memcpy(&n, in.data(), 4);
            if (in.size() != 4 + (size_t)n * 144) die("input size does not match its header");
            PeaksFn f = exe.fn<PeaksFn>("_ZNK4sift4Sift10_findPeaksERKSt5arrayIfLm36EE");
            RSift* s = make_sift(3, 3, 1.6f, 1.4142135f, false);
            for (int j = 0; j < n; ++j) {
                std::array<float, 36> h;
                std::memcpy(h.data(), in.data() + 4 + (size_t)j * 144, 144);
                std::set<float>* r = new std::set<float>(f(s, h));
                put<int32_t>(out, (int32_t)r->size());
                for (float x : *r) put<float>(out, x);
            }
        } else {
            die("unknown command", cmd.c_str());
        }
    } catch (const std::exception& e) {
        // what the reference surfaces as vigra::PreconditionViolation (a std::exception): reported, not a failure of the helper
        std::fclose(out);
        out = std::fopen(argv[4], "wb");
        put<int32_t>(out, -1);
        const std::string w = e.what();
        put(out, w.data(), w.size());
        std::fclose(out);
        return 3;
    }
    std::fclose(out);
    return 0;
}
