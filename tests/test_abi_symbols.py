"""The C-ABI library must load without a GPU and export every entry point include/*.h declares."""
import ctypes, glob, os, re
from av1_base_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = open(h).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        for m in re.finditer(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(av1b_\w+)\s*\(", txt, flags=re.M):
            if "typedef" not in txt[max(0, m.start() - 10):m.start() + 8]:
                names.append(m.group(1))
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = abi.lib()
    syms = declared_symbols()
    assert len(syms) >= 10, syms
    for s in syms:
        assert hasattr(lib, s), "libav1b200.so does not export %s" % s


def test_version_and_no_device_behaviour():
    lib = abi.lib()
    buf = ctypes.create_string_buffer(128)
    assert lib.av1b_version(buf, 128) == 0 and b"av1b200" in buf.value
    if lib.av1b_device_count() == 0:
        cfg = abi.Config()
        lib.av1b_config_default(ctypes.byref(cfg))
        cfg.width, cfg.height = 64, 64
        h = ctypes.c_void_p()
        rc = lib.av1b_encoder_create(ctypes.byref(cfg), ctypes.byref(h))
        assert rc == -2, "without a CUDA device the encoder must fail loudly (no CPU fallback)"
        assert b"no CUDA device" in lib.av1b_last_error()
