import ctypes, numpy as np, sys
hs = ctypes.CDLL('build/libhostsim.so'); orc = ctypes.CDLL('oracle/liborc.so')
P=2**31-1
rng=np.random.default_rng(1)
n=20000
st=rng.integers(0,P,size=(n,16),dtype=np.uint32)
st[0]=np.arange(16); st[1]=0; st[2]=P-1; st[3,:8]=P-1; st[3,8:]=0
a=st.copy(); b=st.copy()
hs.hs_poseidon2_permute(a.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n))
orc.orc_poseidon2_permute_batch(b.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n))
print("permute match:", np.array_equal(a,b), a[0][:4])
