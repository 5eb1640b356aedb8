"""TEST INFRASTRUCTURE ONLY — ctypes front end of the CPU oracle (oracle/_ref/libfastsmc_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package fastsmc_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libfastsmc_oracle.so")


class _Params(C.Structure):
    _fields_ = [
        ("inFileRoot", C.c_char_p), ("decodingQuantFile", C.c_char_p), ("outFileRoot", C.c_char_p),
        ("jobs", C.c_int), ("jobInd", C.c_int), ("foldData", C.c_int), ("usingCSFS", C.c_int),
        ("skipCSFSdistance", C.c_float), ("batchSize", C.c_int), ("skip", C.c_float), ("gap", C.c_int),
        ("max_seeds", C.c_int), ("min_m", C.c_float), ("hashing", C.c_int), ("FastSMC", C.c_int),
        ("BIN_OUT", C.c_int), ("useKnownSeed", C.c_int), ("outputIbdSegmentLength", C.c_int),
        ("time", C.c_int), ("noConditionalAgeEstimates", C.c_int), ("doPerPairPosteriorMean", C.c_int),
        ("doPerPairMAP", C.c_int), ("withinOnly", C.c_int), ("shuffleFlavor", C.c_int), ("simdFlavor", C.c_int),
        ("asmcMode", C.c_int),
    ]


def build(force=False):
    """Compile the oracle with its Makefile (g++ + zlib only)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.fo_create.restype = C.c_void_p
        L.fo_create.argtypes = [C.POINTER(_Params)]
        L.fo_destroy.argtypes = [C.c_void_p]
        L.fo_last_error.restype = C.c_char_p
        L.fo_info.restype = C.c_long
        L.fo_info.argtypes = [C.c_void_p, C.c_int]
        L.fo_probability_threshold.restype = C.c_float
        L.fo_probability_threshold.argtypes = [C.c_void_p]
        L.fo_run.restype = C.c_long
        L.fo_run.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.fo_seed.restype = C.c_long
        L.fo_seed.argtypes = [C.c_void_p]
        for f in ("fo_num_candidates", "fo_num_batches", "fo_num_segments"):
            getattr(L, f).restype = C.c_long
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("fo_last_decode_seconds", "fo_last_pair_sites"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_void_p]
        L.fo_get_transition.restype = C.c_int
        L.fo_get_transition.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        L.fo_round_morgans.restype = C.c_float
        L.fo_round_morgans.argtypes = [C.c_float, C.c_int, C.c_float]
        L.fo_round_physical.restype = C.c_int
        L.fo_round_physical.argtypes = [C.c_int, C.c_int]
        L.fo_get_from_position.restype = C.c_uint
        L.fo_get_from_position.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_float]
        L.fo_get_to_position.restype = C.c_uint
        L.fo_get_to_position.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_float]
        L.fo_cm_between.restype = C.c_double
        L.fo_cm_between.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint, C.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """One loaded dataset + model (reference: Data + HMM construction)."""

    def __init__(self, in_file_root, decoding_quant_file, out_file_root="", *, jobs=1, jobInd=1, foldData=True,
                 usingCSFS=True, skipCSFSdistance=0.0, batchSize=32, skip=0.0, gap=1, max_seeds=0, min_m=1.0,
                 hashing=True, FastSMC=True, BIN_OUT=False, useKnownSeed=True, outputIbdSegmentLength=True,
                 time=100, noConditionalAgeEstimates=False, doPerPairPosteriorMean=False, doPerPairMAP=False,
                 withinOnly=False, shuffleFlavor=0, simdFlavor=False, asmcMode=False):
        L = lib()
        p = _Params(in_file_root.encode(), decoding_quant_file.encode(), out_file_root.encode(), jobs, jobInd,
                    int(foldData), int(usingCSFS), skipCSFSdistance, batchSize, skip, gap, max_seeds, min_m,
                    int(hashing), int(FastSMC), int(BIN_OUT), int(useKnownSeed), int(outputIbdSegmentLength), time,
                    int(noConditionalAgeEstimates), int(doPerPairPosteriorMean), int(doPerPairMAP), int(withinOnly),
                    int(shuffleFlavor), int(simdFlavor), int(asmcMode))
        self._h = L.fo_create(C.byref(p))
        if not self._h:
            raise RuntimeError("oracle: " + L.fo_last_error().decode())
        self.sites = L.fo_info(self._h, 0)
        self.states = L.fo_info(self._h, 1)
        self.num_haps = L.fo_info(self._h, 2)
        self.state_threshold = L.fo_info(self._h, 3)
        self.age_threshold = L.fo_info(self._h, 4)
        self.chr = L.fo_info(self._h, 5)
        self.window_size = L.fo_info(self._h, 6)
        self.w_i = L.fo_info(self._h, 7)
        self.w_j = L.fo_info(self._h, 8)
        self.above_diag = bool(L.fo_info(self._h, 9))
        self.total_samples = L.fo_info(self._h, 10)
        self.csfs_samples = L.fo_info(self._h, 11)
        self.probability_threshold = float(L.fo_probability_threshold(self._h))

    def close(self):
        if self._h:
            lib().fo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- model / data views
    def emissions(self):
        shape = (self.sites, self.states)
        e1, e0m1, e2m0 = (np.empty(shape, np.float32) for _ in range(3))
        lib().fo_get_emissions(C.c_void_p(self._h), _ptr(e1), _ptr(e0m1), _ptr(e2m0))
        return e1, e0m1, e2m0

    def undistinguished(self):
        out = np.empty((self.sites, 3), np.int32)
        lib().fo_get_undistinguished(C.c_void_p(self._h), _ptr(out))
        return out

    def positions(self):
        gen = np.empty(self.sites, np.float32)
        phys = np.empty(self.sites, np.int32)
        lib().fo_get_positions(C.c_void_p(self._h), _ptr(gen), _ptr(phys))
        return gen, phys

    def haplotypes(self):
        out = np.empty((self.num_haps, self.sites), np.uint8)
        for h in range(self.num_haps):
            lib().fo_get_hap(C.c_void_p(self._h), h, _ptr(out[h]))
        return out

    def flipped(self):
        out = np.empty(self.sites, np.uint8)
        lib().fo_get_flipped(C.c_void_p(self._h), _ptr(out))
        return out

    def vector(self, name):
        which = {"initialStateProb": 0, "expectedTimes": 1, "columnRatios": 2, "discretization": 3}[name]
        out = np.empty(self.states + (1 if which == 3 else 0), np.float32)
        lib().fo_get_vector(C.c_void_p(self._h), which, _ptr(out))
        return out

    def transition(self, key):
        out = np.empty((4, self.states), np.float32)
        if lib().fo_get_transition(C.c_void_p(self._h), float(key), _ptr(out)) != 0:
            raise KeyError(key)
        return out  # rows D, B, U, RR

    # ---- drivers
    def run(self, out_path=None, threads=0):
        n = lib().fo_run(C.c_void_p(self._h), out_path.encode() if out_path else None, threads)
        if n < 0:
            raise RuntimeError("oracle: " + lib().fo_last_error().decode())
        return n

    def seed(self):
        n = lib().fo_seed(C.c_void_p(self._h))
        if n < 0:
            raise RuntimeError("oracle: " + lib().fo_last_error().decode())
        return self.candidates()

    def candidates(self):
        n = lib().fo_num_candidates(C.c_void_p(self._h))
        out = np.empty((n, 4), np.uint32)
        if n:
            lib().fo_get_candidates(C.c_void_p(self._h), _ptr(out))
        return out  # hapA, hapB, from, to in decodeFromHashing call order

    def batches(self):
        n = lib().fo_num_batches(C.c_void_p(self._h))
        out = np.empty((n, 5), np.uint32)
        if n:
            lib().fo_get_batches(C.c_void_p(self._h), _ptr(out))
        return out  # nPairs, scanFrom, scanTo, from, to

    def segments(self):
        n = lib().fo_num_segments(C.c_void_p(self._h))
        ints = np.empty((n, 7), np.int32)
        floats = np.empty((n, 3), np.float32)
        if n:
            lib().fo_get_segments(C.c_void_p(self._h), _ptr(ints), _ptr(floats))
        return ints, floats  # (batch,lane,hapA,hapB,posStart,posEnd,mapState), (prob,postMean,map)

    @property
    def last_decode_seconds(self):
        return lib().fo_last_decode_seconds(C.c_void_p(self._h))

    @property
    def last_pair_sites(self):
        return lib().fo_last_pair_sites(C.c_void_p(self._h))

    def decode_posterior(self, hap_a, hap_b, frm=0, to=None):
        to = self.sites if to is None else to
        a = np.ascontiguousarray(hap_a, np.uint32)
        b = np.ascontiguousarray(hap_b, np.uint32)
        out = np.empty((len(a), to - frm, self.states), np.float32)
        rc = lib().fo_decode_posterior(C.c_void_p(self._h), len(a), _ptr(a), _ptr(b), C.c_uint(frm), C.c_uint(to),
                                       _ptr(out))
        if rc != 0:
            raise RuntimeError("oracle: " + lib().fo_last_error().decode())
        return out

    def decode_summary(self, hap_a, hap_b, frm=0, to=None, mean=True, map_=True, ibd=True):
        to = self.sites if to is None else to
        a = np.ascontiguousarray(hap_a, np.uint32)
        b = np.ascontiguousarray(hap_b, np.uint32)
        shape = (len(a), to - frm)
        m = np.empty(shape, np.float32) if mean else None
        mp = np.empty(shape, np.int32) if map_ else None
        ib = np.empty(shape, np.float32) if ibd else None
        rc = lib().fo_decode_summary(C.c_void_p(self._h), len(a), _ptr(a), _ptr(b), C.c_uint(frm), C.c_uint(to),
                                     _ptr(m), _ptr(mp), _ptr(ib))
        if rc != 0:
            raise RuntimeError("oracle: " + lib().fo_last_error().decode())
        return m, mp, ib


def round_morgans(v, precision=2, minv=1e-10):
    return float(lib().fo_round_morgans(v, precision, minv))


def round_physical(v, precision=2):
    return int(lib().fo_round_physical(v, precision))


def get_from_position(gen, frm, cm=0.5):
    g = np.ascontiguousarray(gen, np.float32)
    return int(lib().fo_get_from_position(_ptr(g), len(g), frm, cm))


def get_to_position(gen, to, cm=0.5):
    g = np.ascontiguousarray(gen, np.float32)
    return int(lib().fo_get_to_position(_ptr(g), len(g), to, cm))


def cm_between(w1, w2, gen, word_size=64):
    g = np.ascontiguousarray(gen, np.float32)
    return float(lib().fo_cm_between(w1, w2, _ptr(g), len(g), word_size))


# ---- the reference's own sources, compiled unmodified against oracle/shim (oracle/Makefile: ref_fastsmc_avx / _nosse) ----

def reference_binary(flavour="avx"):
    """Path of the reference build of that SIMD flavour, or None when it has not been built (it is built by `make` in this
    directory wherever /root/reference exists; the GPU box receives the prebuilt file with the snapshot)."""
    path = os.path.join(_HERE, "_ref", f"ref_fastsmc_{flavour}")
    return path if os.access(path, os.X_OK) else None


def reference_command(flavour, in_file_root, decoding_quant_file, out_file_root, **options):
    """argv of one reference run: ASMC::FastSMC(params).run() with the reference's DecodingParams fields set from
    `options` (hashing, jobs, jobInd, time, min_m, skip, gap, max_seeds, batchSize, noConditionalAgeEstimates, bin, ...)."""
    exe = reference_binary(flavour)
    if exe is None:
        raise FileNotFoundError(f"oracle/_ref/ref_fastsmc_{flavour} is not built")
    argv = [exe, f"in={in_file_root}", f"dq={decoding_quant_file}", f"out={out_file_root}"]
    for k, v in options.items():
        argv.append(f"{k}={int(v) if isinstance(v, bool) else v}")
    return argv


def reference_run(flavour, in_file_root, decoding_quant_file, out_file_root, **options):
    """Run the reference build once; returns (timings dict, stderr text).  The output file is
    <out_file_root>.<jobInd>.<jobs>.FastSMC.ibd.gz, named by the reference itself."""
    import json
    r = subprocess.run(reference_command(flavour, in_file_root, decoding_quant_file, out_file_root, **options),
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference build failed ({r.returncode}): {r.stderr[-400:]}")
    return json.loads(r.stdout.strip().splitlines()[-1]), r.stderr
