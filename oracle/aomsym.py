"""Test infrastructure (NOT product code): resolve *local* symbols of the bundled libaom 3.13.1
shared object (OpenCV wheel) through its ELF .symtab so that libaom's own C reference kernels
(av1_inv_txfm2d_add_*_c, aom_lpf_*_c, cdef_filter_*_c, ...) and constant tables can be used as the
oracle's anchor.  Nothing under av1_base_b200/ may import this module.
"""
import ctypes, glob, os, struct, sys

def _site_packages():
    import numpy
    return os.path.dirname(os.path.dirname(numpy.__file__))

def find_lib(pattern):
    hits = sorted(glob.glob(os.path.join(_site_packages(), pattern)))
    if not hits:
        raise FileNotFoundError(pattern)
    return hits[0]

LIBAOM_GLOB = "opencv_python_headless.libs/libaom-*.so*"
LIBAVIF_GLOB = "pillow.libs/libavif-*.so*"

class ElfSyms:
    """Minimal ELF64 .symtab reader: name -> list of (value, size, type)."""
    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            self.data = f.read()
        d = self.data
        assert d[:4] == b"\x7fELF" and d[4] == 2
        shoff = struct.unpack_from("<Q", d, 0x28)[0]
        shentsize, shnum, shstrndx = struct.unpack_from("<HHH", d, 0x3A)
        self.sections = []
        for i in range(shnum):
            name, typ, flags, addr, off, size, link, info, align, entsize = struct.unpack_from(
                "<IIQQQQIIQQ", d, shoff + i * shentsize)
            self.sections.append(dict(name=name, type=typ, addr=addr, off=off, size=size, link=link,
                                      entsize=entsize))
        self.syms = {}
        for s in self.sections:
            if s["type"] != 2:  # SHT_SYMTAB
                continue
            strtab = self.sections[s["link"]]
            n = s["size"] // 24
            for i in range(n):
                st_name, st_info, st_other, st_shndx, st_value, st_size = struct.unpack_from(
                    "<IBBHQQ", d, s["off"] + i * 24)
                if st_shndx == 0 or st_name == 0:
                    continue
                end = d.index(b"\0", strtab["off"] + st_name)
                nm = d[strtab["off"] + st_name:end].decode()
                self.syms.setdefault(nm, []).append((st_value, st_size, st_info & 15, st_shndx))

    def vaddr_to_off(self, vaddr):
        for s in self.sections:
            if s["addr"] and s["type"] != 8 and s["addr"] <= vaddr < s["addr"] + s["size"]:
                return s["off"] + (vaddr - s["addr"])
        raise KeyError(hex(vaddr))

    def read(self, name, which=0):
        """Raw bytes of a data symbol as stored in the file (read-only tables)."""
        v, sz, _, _ = self.syms[name][which]
        off = self.vaddr_to_off(v)
        return self.data[off:off + sz]

_loaded = {}

def load(pattern=LIBAOM_GLOB):
    """Returns (CDLL, base_address, ElfSyms)."""
    if pattern in _loaded:
        return _loaded[pattern]
    path = find_lib(pattern)
    lib = ctypes.CDLL(path)
    # base address: take an exported symbol's runtime address minus its st_value
    es = ElfSyms(path)
    base = None
    for probe in ("aom_codec_version", "aom_codec_av1_dx", "dav1d_version"):
        if probe in es.syms:
            try:
                fn = getattr(lib, probe)
            except AttributeError:
                continue
            base = ctypes.cast(fn, ctypes.c_void_p).value - es.syms[probe][0][0]
            break
    assert base is not None
    _loaded[pattern] = (lib, base, es)
    # libaom's *_c functions call a few helpers through run-time CPU dispatch pointers; fill them
    for init in ("av1_rtcd", "aom_dsp_rtcd", "aom_scale_rtcd"):
        if init in es.syms:
            ctypes.CFUNCTYPE(None)(base + es.syms[init][0][0])()
    return _loaded[pattern]

def func(name, restype, argtypes, pattern=LIBAOM_GLOB, which=0):
    lib, base, es = load(pattern)
    v = es.syms[name][which][0]
    proto = ctypes.CFUNCTYPE(restype, *argtypes)
    return proto(base + v)

if __name__ == "__main__":
    lib, base, es = load()
    print(hex(base), len(es.syms))
    for n in sys.argv[1:]:
        print(n, es.syms.get(n))
