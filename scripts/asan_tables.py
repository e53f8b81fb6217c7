import ctypes as C, numpy as np, sys, os, tempfile
L = C.CDLL(os.environ.get("DCCM_ASAN_LIB", "/tmp/libtables_asan.so"))
f64p = C.POINTER(C.c_double); i32p = C.POINTER(C.c_int32); vp = C.c_void_p
dp = lambda a: a.ctypes.data_as(f64p); ip = lambda a: a.ctypes.data_as(i32p)
L.dccm_table_size.restype = C.c_int64
L.dccm_table_size.argtypes=[vp]; L.dccm_table_free.argtypes=[vp]
L.dccm_last_error.restype = C.c_char_p
def grid(im, jm, gauss=True):
    a = [np.zeros(im), np.zeros(jm), np.zeros(im), np.zeros(jm)]
    fn = L.dccm_grid_gauss if gauss else L.dccm_grid_regular
    assert fn(im, jm, *[dp(x) for x in a]) == 0
    return (im, jm, *a)
def exch(A, O):
    n = A[1] + O[1]
    lat, wt = np.zeros(n), np.zeros(n); jms = C.c_int(0)
    assert L.dccm_grid_exchange(A[1], dp(A[3]), dp(A[5]), O[1], dp(O[5]), C.byref(jms), dp(lat), dp(wt)) == 0
    j = jms.value
    return (A[0], j, A[2].copy(), lat[:j].copy(), A[4].copy(), wt[:j].copy())
def use(h):
    n = L.dccm_table_size(h)
    a = [np.zeros(n, np.int32) for _ in range(4)] + [np.zeros(n)]
    assert L.dccm_table_get(h, *[ip(x) for x in a[:4]], dp(a[4])) == 0
    s, r, c = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n)
    d = tempfile.mkdtemp()
    for wr, rd in ((L.dccm_table_write_text, L.dccm_table_read_text), (L.dccm_table_write_bin, L.dccm_table_read_bin)):
        fn = os.path.join(d, "t").encode()
        assert wr(h, fn) == 0
        h2 = vp(); assert rd(fn, C.byref(h2)) == 0
        assert L.dccm_table_size(h2) == n
        L.dccm_table_free(h2)
    L.dccm_table_free(h)
    return n
tot = 0
for (ia, ja, io, jo, reg) in [(64,32,1,64,False),(128,64,128,64,False),(64,32,72,36,True),(320,160,360,180,True),(16,8,1,8,False),(2,2,2,2,False)]:
    A = grid(ia, ja); O = grid(io, jo, not reg); S = exch(A, O)
    for s, d in [(A,S),(S,A),(S,O),(O,S),(A,O),(O,A)]:
        for order in (1, 2):
            for lm in (0, 1):
                h = vp()
                rc = L.dccm_table_gen_jones99(s[0], dp(s[2]), s[1], dp(s[3]), d[0], dp(d[2]), d[1], dp(d[3]), dp(s[5]), dp(d[5]), order, lm, C.byref(h))
                if rc == 0: tot += use(h)
                # rows variant: three bands
                for j0, j1 in ((1, max(1, d[1]//3)), (d[1]//3 + 1, d[1])):
                    h = vp()
                    rc = L.dccm_table_gen_jones99_rows(s[0], dp(s[2]), s[1], dp(s[3]), d[0], dp(d[2]), d[1], dp(d[3]), dp(s[5]), dp(d[5]), order, lm, j0, j1, C.byref(h))
                    if rc == 0: tot += use(h)
        for lm in (0, 1):
            h = vp()
            rc = L.dccm_table_gen_bilinear(s[0], dp(s[2]), s[1], dp(s[3]), d[0], dp(d[2]), d[1], dp(d[3]), lm, C.byref(h))
            if rc == 0: tot += use(h)
            for fn in (L.dccm_table_gen_bilinear_separable,):
                h = vp()
                rc = fn(s[0], dp(s[2]), s[1], dp(s[3]), d[0], dp(d[2]), d[1], dp(d[3]), lm, C.byref(h))
                if rc == 0: tot += use(h)
            h = vp()
            rc = L.dccm_table_gen_jones99_separable(s[0], dp(s[2]), s[1], dp(s[3]), d[0], dp(d[2]), d[1], dp(d[3]), dp(s[5]), dp(d[5]), 1, lm, C.byref(h))
            if rc == 0: tot += use(h)
for sz in [(8,5,4,3),(360,181,128,65),(1,2,5,3),(5,3,1,2)]:
    h = vp(); assert L.dccm_table_gen_make_mapping_table(*sz, C.byref(h)) == 0; tot += use(h)
# malformed text files
d = tempfile.mkdtemp(); fn = os.path.join(d, "bad.txt")
open(fn, "w").write("1 2 3\n\n  4,5,6,7,1.5D0\nxyz\n1 1 1 1 0.5 extra stuff here that is long " + "x"*600 + "\n2 2 2 2 1e-3")
h = vp(); assert L.dccm_table_read_text(fn.encode(), C.byref(h)) == 0; print("parsed", L.dccm_table_size(h)); L.dccm_table_free(h)
open(fn, "wb").write(b"DCCMTBL1" + b"\x05\x00\x00\x00\x00\x00\x00\x00" + b"\x00" * 10)
h = vp(); print("truncated bin rc", L.dccm_table_read_bin(fn.encode(), C.byref(h)), L.dccm_last_error())
print("entries generated:", tot)
