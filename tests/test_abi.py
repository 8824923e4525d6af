"""Boundary checks that need no GPU: every symbol the public headers declare is exported by the
built library, and the legacy headers still compile as C99 (reference: test/ctest.c, makefile:2)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def declared_functions():
    names = []
    for dirpath, _, files in os.walk(INC):
        for f in files:
            if not f.endswith(".h"):
                continue
            text = open(os.path.join(dirpath, f)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            for m in re.finditer(r"^\s*(?:WORLD_API\s+)?(?:const\s+)?(?:unsigned\s+long\s+long|void|int|double|char)\s*\*?\s*(\w+)\s*\(", text, flags=re.M):
                names.append(m.group(1))
    return sorted(set(names))


def test_headers_declare_the_reference_api():
    names = declared_functions()
    for required in ["Dio", "Harvest", "StoneMask", "CheapTrick", "D4C", "InitializeDioOption",
                     "InitializeHarvestOption", "InitializeCheapTrickOption", "InitializeD4COption",
                     "GetSamplesForDIO", "GetSamplesForHarvest", "GetFFTSizeForCheapTrick",
                     "GetF0FloorForCheapTrick", "Synthesis", "world_b200_synthesis_batch", "world_b200_cheaptrick_batch", "world_b200_d4c_batch",
                     "world_b200_dio_batch", "world_b200_harvest_batch", "world_b200_stonemask_batch",
                     "world_b200_analyze_host", "GetNumberOfAperiodicities", "CodeAperiodicity", "DecodeAperiodicity",
                     "CodeSpectralEnvelope", "DecodeSpectralEnvelope", "world_b200_code_spectral_envelope_batch",
                     "world_b200_code_aperiodicity_batch", "world_b200_decode_spectral_envelope_batch",
                     "world_b200_decode_aperiodicity_batch", "world_b200_pcm_to_double_batch", "world_b200_wav_parse",
                     "world_b200_analyze_coded_host", "wavread", "wavwrite", "WriteF0", "ReadSpectralEnvelope",
                     "world_b200_write_rows", "world_b200_trim"]:
        assert required in names


def test_library_exports_every_declared_symbol():
    from world_b200 import api
    if not os.path.exists(api.DEFAULT_LIB):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(api.DEFAULT_LIB)  # loads without a GPU (cudart is only called on create)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    # and the ctypes table used by the Python mirror covers them all
    assert set(declared_functions()) <= set(api.ABI)


def test_no_gpu_means_loud_failure():
    """The product has no CPU path: creating a context without a device must fail."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from world_b200 import api
    with pytest.raises(api.WorldError):
        api.World(device=0)


def test_headers_are_c99(tmp_path):
    src = tmp_path / "c99.c"
    src.write_text('#include "world_b200.h"\n#include "world/matlabfunctions.h"\n#include "tools/audioio.h"\n'
                   '#include "tools/parameterio.h"\n'
                   'int main(void){DioOption o; RandnState r; InitializeDioOption(&o); randn_reseed(&r); return (int)o.speed;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", INC, str(src)])


def test_option_defaults_and_sizes_match_reference(ref):
    from world_b200 import api
    lib = api.load_library()
    for fs in (8000, 16000, 22050, 44100, 48000):
        a = api.CheapTrickOption(); lib.InitializeCheapTrickOption(fs, ctypes.byref(a))
        b = ref.cheaptrick_option(fs)
        assert (a.q1, a.f0_floor, a.fft_size) == (b.q1, b.f0_floor, b.fft_size)
        assert lib.GetF0FloorForCheapTrick(fs, a.fft_size) == ref.lib.GetF0FloorForCheapTrick(fs, b.fft_size)
        for n in (1, 7, 17500, 160000, 1440000):
            for fp in (1.0, 5.0, 5.5, 10.0):
                assert lib.GetSamplesForDIO(fs, n, fp) == ref.lib.GetSamplesForDIO(fs, n, fp)
                assert lib.GetSamplesForHarvest(fs, n, fp) == ref.lib.GetSamplesForHarvest(fs, n, fp)
                assert lib.world_b200_frames(fs, n, fp) == ref.lib.GetSamplesForDIO(fs, n, fp)
    d = api.DioOption(); lib.InitializeDioOption(ctypes.byref(d)); r = ref.dio_option()
    assert [getattr(d, f) for f, _ in d._fields_] == [getattr(r, f) for f, _ in r._fields_]
    hopt = api.HarvestOption(); lib.InitializeHarvestOption(ctypes.byref(hopt)); r = ref.harvest_option()
    assert [getattr(hopt, f) for f, _ in hopt._fields_] == [getattr(r, f) for f, _ in r._fields_]
    o = api.D4COption(); lib.InitializeD4COption(ctypes.byref(o))
    assert o.threshold == ref.d4c_option().threshold


CUDA_INC = "/usr/local/cuda/include"
CUDA_LIB = "/usr/local/cuda/lib64"


def build_cpp_overload_program(out):
    from world_b200 import api
    libdir = os.path.dirname(api.DEFAULT_LIB)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", INC, "-I", CUDA_INC,
                           os.path.join(ROOT, "tests", "cpp", "batch_overloads.cpp"), "-o", str(out),
                           "-L", libdir, "-lworld_b200", "-L", CUDA_LIB, "-lcudart",
                           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{CUDA_LIB}"])
    return str(out)


def test_cpp_batched_overloads_compile_and_fail_loudly_without_gpu(tmp_path):
    """include/world_b200.hpp (the 'N waveforms at once' overloads of the reference's C API) compiles
    next to the legacy headers, links against the library, and without a device reports an error
    instead of computing anything on the CPU."""
    exe = build_cpp_overload_program(tmp_path / "batch_overloads")
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU path" in r.stderr
