#!/usr/bin/env python3
"""Host-side text I/O and plan code under AddressSanitizer + UBSan (no GPU):

  g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -shared -Iinclude \
      -o /tmp/asan/libhost_asan.so cuda_pro_cell_b200/csrc/hostio.cpp cuda_pro_cell_b200/csrc/textio.cpp
  LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" \
      ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=halt_on_error=1 python tools/host_asan_fuzz.py

30 000 random / token-salad strings go through procell_parse_histogram and procell_parse_cell_types from buffers of the
exact size (no terminator, so a read past the end is caught); every histogram that parses is turned into plans at three
phi values, exported, merged and (sampled) written.  Last run: profiles/r1i_host_sanitizers.txt."""
import ctypes as C, random, os, sys, tempfile
import numpy as np
L = C.CDLL(os.environ.get("PROCELL_ASAN_LIB", "/tmp/asan/libhost_asan.so"))
f64p, u64p, i64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_int64)
class CT(C.Structure): _fields_=[("p",C.c_double),("m",C.c_double),("s",C.c_double)]
L.procell_parse_histogram.argtypes=[C.c_char_p,C.c_size_t,C.POINTER(f64p),C.POINTER(u64p),C.POINTER(C.c_size_t)]
L.procell_parse_cell_types.argtypes=[C.c_char_p,C.c_size_t,C.POINTER(C.POINTER(CT)),C.POINTER(C.c_size_t)]
L.procell_free.argtypes=[C.c_void_p]
L.procell_plan_create.argtypes=[f64p,u64p,C.c_size_t,C.c_double,C.POINTER(C.c_void_p)]
L.procell_plan_destroy.argtypes=[C.c_void_p]
for n in ("n_bins","n_keys","n_rows"): getattr(L,"procell_plan_"+n).restype=C.c_size_t; getattr(L,"procell_plan_"+n).argtypes=[C.c_void_p]
L.procell_plan_export.argtypes=[C.c_void_p,f64p,C.POINTER(C.c_uint32),C.POINTER(C.c_uint32),C.POINTER(C.c_uint8)]
L.procell_merge_rows.argtypes=[C.c_void_p,i64p,C.c_size_t,i64p,i64p]
L.procell_write_histogram.argtypes=[C.c_char_p,C.c_int,C.c_size_t,C.c_size_t,f64p,i64p,i64p]
rng=random.Random(1)
alphabet="0123456789.eE+- \t\n\r-xinfnaINFNAN,;"
tokens=["1","0","12.5","1e3","-4","+7",".5","5.","1e","e5","inf","nan","0x10","1e400","1e-400","18446744073709551615","18446744073709551616","99999999999999999999999","-0","--1","1.2.3","\n","\t"," ","  ","\r\n","4.9e-324","1.7976931348623157e308","abc",""]
def rand_text():
    if rng.random()<0.5:
        return "".join(rng.choice(alphabet) for _ in range(rng.randrange(0,200)))
    return "".join(rng.choice(tokens)+rng.choice([" ","\n","\t",""," \n"]) for _ in range(rng.randrange(0,60)))
n_plans=0
for it in range(30000):
    t=rand_text().encode()
    # exact-size buffer (no terminator) so that ASan sees any read past the end
    buf=(C.c_char*max(len(t),1)).from_buffer_copy(t if t else b"\0")
    v=f64p();f=u64p();n=C.c_size_t()
    rc=L.procell_parse_histogram(C.cast(buf,C.c_char_p),len(t),C.byref(v),C.byref(f),C.byref(n))
    if rc==0 and n.value:
        vals=np.ctypeslib.as_array(v,(n.value,)).copy(); fr=np.ctypeslib.as_array(f,(n.value,)).copy()
        for phi in (0.0, 0.5, 1e-9):
            h=C.c_void_p()
            if L.procell_plan_create(vals.ctypes.data_as(f64p),fr.ctypes.data_as(u64p),n.value,phi,C.byref(h))==0:
                n_plans+=1
                nb,nk,nr=L.procell_plan_n_bins(h),L.procell_plan_n_keys(h),L.procell_plan_n_rows(h)
                rv=np.zeros(nr+1);kr=np.zeros(nk+1,np.uint32);kb=np.zeros(nb+1,np.uint32);kd=np.zeros(nb+1,np.uint8)
                L.procell_plan_export(h,rv.ctypes.data_as(f64p),kr.ctypes.data_as(C.POINTER(C.c_uint32)),kb.ctypes.data_as(C.POINTER(C.c_uint32)),kd.ctypes.data_as(C.POINTER(C.c_uint8)))
                T=3; cnt=np.ones((nk,T),np.int64); rf=np.zeros(nr+1,np.int64); rr=np.zeros((nr+1)*T,np.int64)
                L.procell_merge_rows(h,cnt.ctypes.data_as(i64p),T,rf.ctypes.data_as(i64p),rr.ctypes.data_as(i64p))
                if it%50==0:
                    with tempfile.NamedTemporaryFile() as tf:
                        L.procell_write_histogram(tf.name.encode(),1,T,nr,rv.ctypes.data_as(f64p),rf.ctypes.data_as(i64p),rr.ctypes.data_as(i64p))
                L.procell_plan_destroy(h)
    if rc==0: L.procell_free(v); L.procell_free(f)
    ct=C.POINTER(CT)();m=C.c_size_t()
    rc=L.procell_parse_cell_types(C.cast(buf,C.c_char_p),len(t),C.byref(ct),C.byref(m))
    if rc==0 or ct: L.procell_free(ct)
print("fuzz done, plans built:",n_plans)
