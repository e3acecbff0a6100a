"""Mutation fuzzing of the host-side file readers (csrc/fast5.cu, csrc/vbz.cu) built with AddressSanitizer + UBSan
(scripts/host_fuzz_asan.sh builds /tmp/libh5asan.so and runs this).  usage: host_fuzz.py SEED N_HDF5 N_ZSTD"""
import ctypes as C, sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import numpy as np, h5_writer as hw
lib=C.CDLL('/tmp/libh5asan.so')
lib.nd_fast5_read_signal.argtypes=[C.c_char_p,C.c_int64,C.POINTER(C.c_int16),C.c_int64,C.POINTER(C.c_int64),C.c_char_p,C.c_int32,C.c_char_p,C.c_int32]
lib.nd_zstd_decompress.argtypes=[C.c_char_p,C.c_int64,C.POINTER(C.c_uint8),C.c_int64,C.POINTER(C.c_int64),C.c_char_p,C.c_int32]
lib.nd_fast5_list_reads.argtypes=[C.c_char_p,C.c_int64,C.c_char_p,C.c_int64,C.POINTER(C.c_int64),C.POINTER(C.c_int32),C.POINTER(C.c_int32),C.c_char_p,C.c_int32]
lib.nd_fast5_read_signal_of.argtypes=[C.c_char_p,C.c_int64,C.c_char_p,C.POINTER(C.c_int16),C.c_int64,C.POINTER(C.c_int64),C.c_char_p,C.c_int32]
def dac(n, seed=0):
    rng = np.random.default_rng(seed)
    levels = np.repeat(rng.integers(350, 700, n // 9 + 1), 9)[:n]
    return (levels + rng.integers(-12, 13, n)).astype(np.int16)
def read(raw, cap=20000):
    out=(C.c_int16*cap)(); cnt=C.c_int64(0); name=C.create_string_buffer(64); err=C.create_string_buffer(256)
    rc=lib.nd_fast5_read_signal(raw,len(raw),out,cap,C.byref(cnt),name,64,err,256)
    return rc
def readof(raw, nm, cap=20000):
    out=(C.c_int16*cap)(); cnt=C.c_int64(0); err=C.create_string_buffer(256)
    return lib.nd_fast5_read_signal_of(raw,len(raw),nm,out,cap,C.byref(cnt),err,256)
lib.nd_h5_list_group.argtypes=[C.c_char_p,C.c_int64,C.c_char_p,C.c_char_p,C.c_int64,C.POINTER(C.c_int64),C.POINTER(C.c_int32),C.c_char_p,C.c_int32]
lib.nd_h5_read_dataset.argtypes=[C.c_char_p,C.c_int64,C.c_char_p,C.POINTER(C.c_uint8),C.c_int64,C.POINTER(C.c_int64),C.c_char_p,C.c_int32]
def generic(raw):
    names=C.create_string_buffer(4096); need=C.c_int64(0); n=C.c_int32(0); err=C.create_string_buffer(256)
    for path in (b"/", b"/Raw/Reads", b"/UniqueGlobalKey"):
        lib.nd_h5_list_group(raw,len(raw),path,names,4096,C.byref(need),C.byref(n),err,256)
    info=(C.c_int64*8)(); out=(C.c_uint8*40000)()
    lib.nd_h5_read_dataset(raw,len(raw),b"/Raw/Reads/Read_17/Signal",out,40000,info,err,256)
def lst(raw):
    names=C.create_string_buffer(4096); need=C.c_int64(0); n=C.c_int32(0); lay=C.c_int32(0); err=C.create_string_buffer(256)
    return lib.nd_fast5_list_reads(raw,len(raw),names,4096,C.byref(need),C.byref(n),C.byref(lay),err,256)
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
sig=dac(6000,9)
files=[hw.make_fast5(sig,chunk=1000,filters=(2,1)), hw.make_fast5(sig,chunk=1500,filters=(32020,),kw_vbz_version=1),
       hw.make_fast5(sig,chunk=1500,filters=(32020,),kw_vbz_version=0), hw.make_fast5(sig,flavour="new",chunk=None),
       hw.make_fast5(sig,flavour="new",chunk="implicit"), hw.make_fast5(sig[:2000],chunk="compact"),
       hw.make_fast5(sig,chunk=700,filters=(3,2,1),other_reads=("Read_1","Read_2"),group_levels=2, fan=3),
       hw.make_multi_fast5({"a":sig[:3000],"b":sig[3000:]},chunk=512)]
ok=bad=0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 3000):
    f=bytearray(files[it%len(files)])
    k=int(rng.integers(1,6))
    mode=int(rng.integers(0,4))
    for _ in range(k):
        pos=int(rng.integers(0,len(f)))
        if mode==0: f[pos]^=1<<int(rng.integers(0,8))
        elif mode==1: f[pos]=int(rng.integers(0,256))
        elif mode==2: f[pos:pos+4]=bytes(rng.integers(0,256,4,dtype=np.uint8))
        else: f[pos:pos+8]=b"\xff"*8
    if rng.random()<0.1: f=f[:int(rng.integers(0,len(f)))]
    f=bytes(f)
    r=read(f); lst(f); readof(f,b"read_a"); readof(f,b"Read_17"); generic(f)
    ok+= r==0; bad+= r!=0
print("hdf5 fuzz: ok",ok,"errors",bad)
# zstd fuzz
text=b" ".join(str(int(x)).encode() for x in dac(20000,2))
blobs=[hw.zstd_compress(text,l) for l in (1,3,19)]+[hw.zstd_compress(dac(30000,3).tobytes(),l) for l in (1,9)]
ok=bad=0
out=(C.c_uint8*400000)(); cnt=C.c_int64(0); err=C.create_string_buffer(256)
for it in range(int(sys.argv[3]) if len(sys.argv)>3 else 4000):
    f=bytearray(blobs[it%len(blobs)])
    for _ in range(int(rng.integers(1,4))):
        pos=int(rng.integers(0,len(f)))
        if rng.random()<0.5: f[pos]^=1<<int(rng.integers(0,8))
        else: f[pos]=int(rng.integers(0,256))
    if rng.random()<0.1: f=f[:int(rng.integers(0,len(f)))]
    r=lib.nd_zstd_decompress(bytes(f),len(f),out,400000,C.byref(cnt),err,256)
    ok+= r==0; bad+= r!=0
print("zstd fuzz: ok",ok,"errors",bad)
