"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product (recursive-stwo_b200/).

CPU restatement of the reference's circuit DSL and of the recursive verifier circuit built with it: every
DSL call computes its value eagerly and appends rows to the Plonk-with-Poseidon constraint system exactly in
the reference's append order, so `variables[]`, the wiring columns and the 22 trace columns are defined
bit-exactly.  Pure Python on purpose (a second language and a second structure next to the C++ recorder of the
product); Poseidon2 permutations go through the C oracle (oracle/liborc.so).

Follows (all paths relative to /root/reference):
  constraint_system/src/plonk_with_poseidon.rs          CS
  primitives/fields/src/{m31,cm31,qm31}.rs              m31_*/cm31_*/qm31_* helpers on V
  primitives/bits/src/lib.rs                            Bits
  primitives/poseidon31/src/lib.rs                      Half / permute
  primitives/merkle/src/lib.rs                          hash_*_columns_*
  primitives/channel/src/lib.rs                         Channel
  primitives/circle/src/lib.rs, query/src/lib.rs, line/src/lib.rs
  components/recursive/{data_structures,fiat_shamir,composition,answer,folding}/src/*.rs
  examples/single-proof/src/main.rs:33-90, examples/multi-proofs/src/main.rs:49-139   verifier_circuit()

Parity status: PINNED by (i) check_arithmetics + check_poseidon_invocations of the reference itself holding on
the produced system, (ii) every in-circuit equalverify holding on accepted fixtures, (iii) the padded circuit
sizes recorded in the header of the NEXT proof of the reference's recursion chain (small_proof -> 16/15,
recursive_proof_16_15 x5 -> 19/18, ... examples/multi-proofs/src/main.rs:173-296), (iv) Fiat-Shamir draws,
answers and folds equal to the C oracle's (itself pinned by SURVEY App. F).  One reference nondeterminism is
fixed by convention: AnswerResults::compute iterates a HashSet of mask shifts (answer/src/lib.rs:45-72); we use
first-appearance order (0 then -1), Plonk set before Poseidon set.
"""
import ctypes
import hashlib
import os
import struct

import numpy as np

P = (1 << 31) - 1
_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _orc():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(os.path.join(_HERE, "liborc.so"))
    return _lib


_ST = ctypes.c_uint32 * 16


def poseidon2_permute(state):
    buf = _ST(*state)
    _orc().orc_poseidon2_permute(buf)
    return list(buf)


# ---- field arithmetic on tuples (stwo M31 / CM31 / QM31) ------------------------------------------------
def m_inv(a):
    return pow(a, P - 2, P)


def c_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def c_inv(a):
    n = m_inv((a[0] * a[0] + a[1] * a[1]) % P)
    return (a[0] * n % P, (P - a[1]) * n % P)


def q_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P, (a[2] + b[2]) % P, (a[3] + b[3]) % P)


def q_mul(a, b):
    a0, a1, b0, b1 = a[0:2], a[2:4], b[0:2], b[2:4]
    ac = c_mul(a0, b0)
    bd = c_mul(a1, b1)
    r = c_mul(bd, (2, 1))
    ad = c_mul(a0, b1)
    bc = c_mul(a1, b0)
    return ((ac[0] + r[0]) % P, (ac[1] + r[1]) % P, (ad[0] + bc[0]) % P, (ad[1] + bc[1]) % P)


def q_scale(a, k):
    return (a[0] * k % P, a[1] * k % P, a[2] * k % P, a[3] * k % P)


def q_inv(a):
    # (a + bu)^-1 = (a - bu) / (a^2 - (2+i) b^2)
    a0, a1 = a[0:2], a[2:4]
    b2 = c_mul(a1, a1)
    t = c_mul(b2, (2, 1))
    a2 = c_mul(a0, a0)
    den = c_inv(((a2[0] - t[0]) % P, (a2[1] - t[1]) % P))
    x = c_mul(a0, den)
    y = c_mul(((P - a1[0]) % P, (P - a1[1]) % P), den)
    return (x[0], x[1], y[0], y[1])


Q0, Q1 = (0, 0, 0, 0), (1, 0, 0, 0)


def qm(v):
    return (v % P, 0, 0, 0)


# ---- circle group over M31 ---------------------------------------------------------------------------
G = (2, 1268011823)


def cp_add(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def cp_double(a):
    return cp_add(a, a)


def cp_neg(a):
    return (a[0], (P - a[1]) % P)


def cp_gen(k):
    g = G
    for _ in range(31 - k):
        g = cp_double(g)
    return g


# ---- constraint system (constraint_system/src/plonk_with_poseidon.rs) ------------------------------------------------
LOG_ADD, LOG_MUL, LOG_MULC, LOG_IN, LOG_INV_M31, LOG_INV_QM31, LOG_CINV_RE, LOG_CINV_IM, LOG_COORD, LOG_BIT, LOG_PERM = range(1, 12)
NO_VAR = 0xFFFFFFFF


class CS:
    kind = "with"                                                        # ConstraintSystemType::PlonkWithPoseidon

    def __init__(self):                                                  # :43-99
        self.variables = [Q0, Q1, (0, 1, 0, 0), (0, 0, 1, 0)]
        self.cache = {}
        self.a_wire, self.b_wire, self.c_wire = [0, 1, 2, 3], [0, 0, 0, 0], [0, 1, 2, 3]
        self.poseidon_wire, self.enforce_c_m31, self.op = [0] * 4, [0] * 4, [1] * 4
        self.flow = []                                                   # (e1, e2, e3, e4, (addr, swap)), e = (wire, hash8)
        self.num_input = 3
        self.mult_a = self.mult_b = self.mult_c = self.mult_poseidon = None
        self.n_perm = 0
        # how every variable got its value, in creation order: [kind, dst, a, b, v0..v3] (orc_tape.c replays it in C as the
        # CPU baseline of the circuit's value arithmetic).  kinds: LOG_*.
        self.log = []
        self.perm_log = []

    def _row(self, a, b, c, op, pw=0, enf=0):
        self.a_wire.append(a); self.b_wire.append(b); self.c_wire.append(c)
        self.poseidon_wire.append(pw); self.enforce_c_m31.append(enf); self.op.append(op % P)

    def insert_gate(self, a, b, c, op):                                  # :101-115
        n = len(self.variables)
        assert a < n and b < n and c < n
        self._row(a, b, c, op)

    def enforce_zero(self, var):                                         # :130-139
        self._row(var, 0, 0, 1)

    def add(self, a, b):                                                 # :141-150
        c = len(self.variables)
        self.variables.append(q_add(self.variables[a], self.variables[b]))
        self.log.append((LOG_ADD, c, a, b, 0, 0, 0, 0))
        self.insert_gate(a, b, c, 1)
        return c

    def mul(self, a, b):                                                 # :173-182
        c = len(self.variables)
        self.variables.append(q_mul(self.variables[a], self.variables[b]))
        self.log.append((LOG_MUL, c, a, b, 0, 0, 0, 0))
        self.insert_gate(a, b, c, 0)
        return c

    def mul_constant(self, a, k):                                        # :184-192
        k %= P
        c = len(self.variables)
        self.variables.append(q_scale(self.variables[a], k))
        self.log.append((LOG_MULC, c, a, k, 0, 0, 0, 0))
        self.insert_gate(a, 0, c, k)
        return c

    def assemble_poseidon_gate(self, a, b):                              # :152-171
        c = len(self.variables)
        self.variables.append(q_mul(self.variables[a], self.variables[b]))
        self.log.append((LOG_MUL, c, a, b, 0, 0, 0, 0))
        self._row(a, b, c, 0, pw=c)
        return c

    def _log_def(self, c, src):
        # src: None = taken from the proof / hint structs, (kind, a, b) = derived from an earlier variable, "perm" = logged by permute
        if src is None:
            self.log.append((LOG_IN, c, 0, 0) + tuple(self.variables[c]))
        elif src != "perm":
            self.log.append((src[0], c, src[1], src[2], 0, 0, 0, 0))

    def new_m31(self, v, mode, src=None):                                # :194-233
        c = len(self.variables)
        self.variables.append(qm(v))
        if mode == "constant":
            self.log.append((LOG_MULC, c, 1, v % P, 0, 0, 0, 0))
        else:
            self._log_def(c, src)
        if mode == "input":
            self._row(c, 0, c, 1, enf=1)
            self.num_input += 1
        elif mode == "witness":
            self._row(c, 0, c, 1, enf=1)
        else:
            self._row(1, 0, c, v)
        return c

    def new_qm31(self, v, mode, src=None):                               # :235-281
        c = len(self.variables)
        self.variables.append(tuple(v))
        if mode != "constant":
            self._log_def(c, src)
        if mode == "input":
            self._row(c, 0, c, 1, enf=1)
            self.num_input += 1
        elif mode == "constant":
            fr = self.new_m31(v[0], "constant")
            fi = self.new_m31(v[1], "constant")
            sr = self.new_m31(v[2], "constant")
            si = self.new_m31(v[3], "constant")
            t = self.mul(fi, 2)
            a = self.add(fr, t)
            t = self.mul(si, 2)
            t = self.add(sr, t)
            b = self.mul(t, 3)
            self.log.append((LOG_ADD, c, a, b, 0, 0, 0, 0))
            self._row(a, b, c, 1)
        return c

    def invoke_poseidon_accelerator(self, e1, e2, e3, e4, swap):
        self.flow.append((e1, e2, e3, e4, swap))

    # ---- finalisation ------------------------------------------------------------------------------------------------
    def pad(self):                                                       # :283-335
        self.n_rows_unpadded, self.n_flow_unpadded = len(self.a_wire), len(self.flow)
        n = len(self.flow)
        self.n_flow_padded = max(32, (n + 15) // 16 * 16)               # padding entries use CONSTANT_1/2/3 (parity unpinned)
        n = len(self.a_wire)
        padded = 1 << (n - 1).bit_length()
        for _ in range(n, padded):
            self._row(0, 0, 0, 1)

    def check_arithmetics(self):                                         # :337-380
        v = self.variables
        for i in range(len(self.a_wire)):
            a, b, c, op = v[self.a_wire[i]], v[self.b_wire[i]], v[self.c_wire[i]], self.op[i]
            want = q_add(q_scale(q_add(a, b), op), q_scale(q_mul(a, b), (1 - op) % P))
            if want != c:
                return i
            if self.enforce_c_m31[i] and (c[1] or c[2] or c[3]):
                return i
        return -1

    def populate_logup_arguments(self):                                  # :382-466
        nv, nr = len(self.variables), len(self.a_wire)
        assert nr & (nr - 1) == 0
        counts = [0] * nv
        for i in range(nr):
            counts[self.a_wire[i]] += 1; counts[self.b_wire[i]] += 1; counts[self.c_wire[i]] += 1
        for i in range(self.num_input):
            counts[i + 1] += 1
        for f in self.flow:
            counts[f[4][0]] += 1
        seen = [False] * nv
        ma, mb, mc = [], [], []
        for i in range(nr):
            for w, out in ((self.a_wire[i], ma), (self.b_wire[i], mb), (self.c_wire[i], mc)):
                if seen[w]:
                    out.append(1)
                else:
                    seen[w] = True
                    out.append(1 - counts[w])
        mpv = [0] * nv
        for f in self.flow:
            for k in range(4):
                mpv[f[k][0]] += 1
        mpv[0] = 0
        mp = []
        for i in range(nr):
            r = mpv[self.poseidon_wire[i]]
            if r:
                assert counts[self.poseidon_wire[i]] == 1
                mpv[self.poseidon_wire[i]] = 0
            mp.append(r)
        self.mult_a, self.mult_b, self.mult_c, self.mult_poseidon = ma, mb, mc, mp

    def check_poseidon_invocations(self):                                # :468-519
        m = {}
        v = self.variables
        for i in range(len(self.a_wire)):
            if self.mult_poseidon[i]:
                m[self.poseidon_wire[i]] = list(v[self.a_wire[i]]) + list(v[self.b_wire[i]])
        for n, (r1, r2, r3, r4, sw) in enumerate(self.flow):
            for r in (r1, r2, r3, r4):
                if r[0] and m[r[0]] != list(r[1]):
                    return n
            st = list(r2[1]) + list(r1[1]) if sw[1] else list(r1[1]) + list(r2[1])
            if poseidon2_permute(st) != list(r3[1]) + list(r4[1]):
                return n
        return -1

    def trace_columns(self):
        """generate_plonk_with_poseidon_circuit (:521-628): uint32 [22, n_rows] in the struct-literal order
        mult_a, mult_b, mult_c, poseidon_wire, mult_poseidon, enforce_c_m31, a_wire, b_wire, c_wire, op,
        a_val_0..3, b_val_0..3, c_val_0..3."""
        n = len(self.a_wire)
        out = np.zeros((22, n), dtype=np.uint32)
        for k, m in enumerate((self.mult_a, self.mult_b, self.mult_c)):
            out[k] = np.array([x % P for x in m], dtype=np.uint32)
        out[3] = self.poseidon_wire; out[4] = self.mult_poseidon; out[5] = self.enforce_c_m31
        out[6] = self.a_wire; out[7] = self.b_wire; out[8] = self.c_wire; out[9] = self.op
        va = np.array(self.variables, dtype=np.uint32)
        out[10:14] = va[np.array(self.a_wire)].T
        out[14:18] = va[np.array(self.b_wire)].T
        out[18:22] = va[np.array(self.c_wire)].T
        return out

    def flow_arrays(self):
        """(wire [n,4], swap_addr [n], hash [n,32], swap [n]) of the un-padded Poseidon flow"""
        n = len(self.flow)
        wire = np.zeros((n, 4), dtype=np.uint32); addr = np.zeros(n, dtype=np.uint32)
        h = np.zeros((n, 32), dtype=np.uint32); sw = np.zeros(n, dtype=np.uint8)
        for i, f in enumerate(self.flow):
            for k in range(4):
                wire[i, k] = f[k][0]; h[i, 8 * k:8 * k + 8] = f[k][1]
            addr[i], sw[i] = f[4][0], 1 if f[4][1] else 0
        return wire, addr, h, sw


def _poseidon2_constants():
    """round constants out of include/stwo_b200_poseidon2_constants.h (= primitives/poseidon31/src/parameters.rs:6-190)"""
    import re
    txt = open(os.path.join(_HERE, "..", "include", "stwo_b200_poseidon2_constants.h")).read()
    out = {}
    for name in ("DIAG16", "RC_FIRST", "RC_PARTIAL", "RC_LAST"):
        body = re.search(r"#define STWO_P2_%s \{(.*?)\}" % name, txt, re.S).group(1)
        out[name] = [int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]+)u", body)]
    return out


_P2 = None


def p2_constants():
    global _P2
    if _P2 is None:
        _P2 = _poseidon2_constants()
    return _P2


LOG_M4, LOG_POW5M4, LOG_HADAMARD, LOG_GRANDSUM, LOG_POW4 = range(12, 17)


def _m4(x):
    t0 = (x[0] + x[1]) % P
    t1 = (x[2] + x[3]) % P
    t2 = (2 * x[1] + t1) % P
    t3 = (2 * x[3] + t0) % P
    t4 = (4 * t1 + t3) % P
    t5 = (4 * t0 + t2) % P
    return ((t3 + t5) % P, t5, (t2 + t4) % P, t4)


def _had(a, b):
    return tuple(a[k] * b[k] % P for k in range(4))


class CSWithout(CS):
    """constraint_system/src/plonk_without_poseidon.rs: rows carry four selectors (op1..op4) and five more gate kinds"""
    kind = "without"

    def __init__(self):                                                  # :33-90
        CS.__init__(self)
        self.op1, self.op2, self.op3, self.op4 = [1] * 4, [0] * 4, [0] * 4, [0] * 4
        self.op = self.op1                                               # the arithmetic selector
        del self.poseidon_wire, self.enforce_c_m31

    def _row(self, a, b, c, op1, op2=0, op3=0, op4=0):
        self.a_wire.append(a); self.b_wire.append(b); self.c_wire.append(c)
        self.op1.append(op1 % P); self.op2.append(op2); self.op3.append(op3); self.op4.append(op4)

    def _gate(self, kind, a, b, value, sel):
        c = len(self.variables)
        self.variables.append(value)
        self.log.append((kind, c, a, b, 0, 0, 0, 0))
        self._row(a, b, c, *sel)
        return c

    def do_m4_gate(self, a, b):                                          # :108-139 (b is not used by the value)
        return self._gate(LOG_M4, a, b, _m4(self.variables[a]), (1, 0, 1, 0))

    def do_pow5m4_gate(self, a, b):                                      # :140-173
        return self._gate(LOG_POW5M4, a, b, _m4(_had(self.variables[a], self.variables[b])), (1, 1, 1, 0))

    def do_pow5_gate(self, a, b):                                        # :174-198
        return self._gate(LOG_HADAMARD, a, b, _had(self.variables[a], self.variables[b]), (1, 1, 0, 1))

    def do_hadamard(self, a, b):                                         # :199-223
        return self._gate(LOG_HADAMARD, a, b, _had(self.variables[a], self.variables[b]), (1, 0, 0, 1))

    def do_grandsum_gate(self, a, b):                                    # :224-245
        s = (sum(self.variables[a]) + sum(self.variables[b])) % P
        return self._gate(LOG_GRANDSUM, a, b, (s, s, s, s), (1, 0, 1, 1))

    def assemble_poseidon_gate(self, a, b):
        raise NotImplementedError("unimplemented!() in the reference (constraint_system/src/lib.rs:267-277)")

    def new_m31(self, v, mode, src=None):                                # :290-334
        c = len(self.variables)
        self.variables.append(qm(v))
        if mode == "constant":
            self.log.append((LOG_MULC, c, 1, v % P, 0, 0, 0, 0))
            self._row(1, 0, c, v)
        else:
            self._log_def(c, src)
            self._row(c, 1, c, 1, 0, 0, 1)                               # hadamard with variable 1 forces an M31
            if mode == "input":
                self.num_input += 1
        return c

    def new_qm31(self, v, mode, src=None):                               # :335-391
        c = len(self.variables)
        self.variables.append(tuple(v))
        if mode == "constant":
            fr, fi = self.new_m31(v[0], "constant"), self.new_m31(v[1], "constant")
            sr, si = self.new_m31(v[2], "constant"), self.new_m31(v[3], "constant")
            t = self.mul(fi, 2)
            a = self.add(fr, t)
            t = self.mul(si, 2)
            t = self.add(sr, t)
            b = self.mul(t, 3)
            self.log.append((LOG_ADD, c, a, b, 0, 0, 0, 0))
            self._row(a, b, c, 1)
        else:
            self._log_def(c, src)
            self._row(c, 0, c, 1)
            if mode == "input":
                self.num_input += 1
        return c

    def pad(self):                                                       # :392-409
        self.n_rows_unpadded, self.n_flow_unpadded, self.n_flow_padded = len(self.a_wire), 0, 0
        n = len(self.a_wire)
        for _ in range(n, 1 << (n - 1).bit_length()):
            self._row(0, 0, 0, 1)

    def check_arithmetics(self):                                         # :410-599
        v = self.variables
        for i in range(len(self.a_wire)):
            a, b, c = v[self.a_wire[i]], v[self.b_wire[i]], v[self.c_wire[i]]
            op1, sel = self.op1[i], (self.op2[i], self.op3[i], self.op4[i])
            pow4 = tuple(pow(x, 4, P) for x in a)
            had = _had(a, b)
            if sel == (0, 0, 0):
                want = q_add(q_scale(q_add(a, b), op1), q_scale(q_mul(a, b), (1 - op1) % P))
            elif sel == (0, 0, 1):
                want = had
            elif sel == (1, 1, 0):
                want = _m4(had)
                if b != pow4:
                    return i
            elif sel == (1, 0, 1):
                want = had
                if b != pow4:
                    return i
            elif sel == (0, 1, 0):
                want = _m4(had)
            elif sel == (0, 1, 1):
                s = (sum(a) + sum(b)) % P
                want = (s, s, s, s)
            else:
                return i
            if sel != (0, 0, 0) and op1 != 1:
                return i
            if want != c:
                return i
        return -1

    def populate_logup_arguments(self):                                  # :600-632
        nv, nr = len(self.variables), len(self.a_wire)
        counts = [0] * nv
        for i in range(nr):
            counts[self.a_wire[i]] += 1; counts[self.b_wire[i]] += 1; counts[self.c_wire[i]] += 1
        for i in range(self.num_input):
            counts[i + 1] += 1
        seen = [False] * nv
        mc = []
        for i in range(nr):
            w = self.c_wire[i]
            if not seen[w]:
                mc.append(-(counts[w] - 1))
                seen[w] = True
            else:
                mc.append(1)
        self.mult_c = mc

    def check_poseidon_invocations(self):
        return -1                                                        # unimplemented!() for this system: nothing to check

    def trace_columns(self):
        """generate_plonk_without_poseidon_circuit (:633-713): uint32 [20, n_rows]: mult_c, a_wire, b_wire, c_wire, op1..op4,
        a_val_0..3, b_val_0..3, c_val_0..3"""
        n = len(self.a_wire)
        out = np.zeros((20, n), dtype=np.uint32)
        out[0] = np.array([x % P for x in self.mult_c], dtype=np.uint32)
        out[1] = self.a_wire; out[2] = self.b_wire; out[3] = self.c_wire
        out[4] = self.op1; out[5] = self.op2; out[6] = self.op3; out[7] = self.op4
        va = np.array(self.variables, dtype=np.uint32)
        out[8:12] = va[np.array(self.a_wire)].T
        out[12:16] = va[np.array(self.b_wire)].T
        out[16:20] = va[np.array(self.c_wire)].T
        return out


# ---- variables: one class, value = QM31 4-tuple; the m31_/cm31_/qm31_ prefix says which reference impl a helper follows ----
class V:
    """rank: 0 = M31Var, 1 = CM31Var, 2 = QM31Var.  The reference implements `low + high` and `low * high` as
    `high + low` / `high * low` (cm31.rs:100-105,160-165; qm31.rs:99-104,121-126,187-192,210-215), which decides the
    a_wire/b_wire order of the row."""
    __slots__ = ("cs", "value", "variable", "rank")

    def __init__(self, cs, value, variable, rank):
        self.cs, self.value, self.variable, self.rank = cs, value, variable, rank

    def __add__(self, o):
        a, b = (o, self) if self.rank < o.rank else (self, o)
        return V(self.cs, q_add(self.value, o.value), self.cs.add(a.variable, b.variable), a.rank)

    def __mul__(self, o):
        a, b = (o, self) if self.rank < o.rank else (self, o)
        return V(self.cs, q_mul(self.value, o.value), self.cs.mul(a.variable, b.variable), a.rank)

    def __neg__(self):                                                   # Neg = mul_constant(-1)
        return V(self.cs, q_scale(self.value, P - 1), self.cs.mul_constant(self.variable, P - 1), self.rank)

    def __sub__(self, o):                                                # Sub = self + &(-rhs) for every type pair
        n = -o
        return self + n

    def mul_constant_m31(self, k):
        return V(self.cs, q_scale(self.value, k % P), self.cs.mul_constant(self.variable, k), self.rank)

    def shift_by_i(self):
        return V(self.cs, q_mul(self.value, (0, 1, 0, 0)), self.cs.mul(self.variable, 2), max(self.rank, 1))

    def shift_by_j(self):
        return V(self.cs, q_mul(self.value, (0, 0, 1, 0)), self.cs.mul(self.variable, 3), 2)

    def shift_by_ij(self):
        return self.shift_by_i().shift_by_j()

    def equalverify(self, o):
        assert self.value == o.value, ("equalverify", self.value, o.value)
        self.cs.insert_gate(self.variable, 0, o.variable, 1)


def m31_zero(cs):
    return V(cs, Q0, 0, 0)


def m31_one(cs):
    return V(cs, Q1, 1, 0)


def qm31_zero(cs):
    return V(cs, Q0, 0, 2)


def qm31_one(cs):
    return V(cs, Q1, 1, 2)


def m31_witness(cs, v, src=None):
    return V(cs, qm(v), cs.new_m31(v % P, "witness", src), 0)


def m31_constant(cs, v):                                                 # m31.rs:33-59
    v %= P
    if v == 0:
        return m31_zero(cs)
    if v == 1:
        return m31_one(cs)
    key = "m31 %d" % v
    if key in cs.cache:
        return V(cs, qm(v), cs.cache[key], 0)
    var = cs.new_m31(v, "constant")
    cs.cache[key] = var
    return V(cs, qm(v), var, 0)


def m31_inv(x):                                                          # m31.rs:136-143
    r = m31_witness(x.cs, m_inv(x.value[0]), (LOG_INV_M31, x.variable, 0))
    x.cs.insert_gate(x.variable, r.variable, 1, 0)
    return r


def cm31_witness(cs, v, src=(None, None)):                               # cm31.rs:27-37
    re = m31_witness(cs, v[0], src[0])
    im = m31_witness(cs, v[1], src[1])
    return V(cs, (v[0] % P, v[1] % P, 0, 0), cs.add(re.variable, cs.mul(im.variable, 2)), 1)


def cm31_inv(x):                                                         # cm31.rs:238-243 (no constraint row!)
    return cm31_witness(x.cs, c_inv(x.value[0:2]), ((LOG_CINV_RE, x.variable, 0), (LOG_CINV_IM, x.variable, 0)))


def qm31_witness(cs, v, src=None):
    return V(cs, tuple(v), cs.new_qm31(v, "witness", src), 2)


def qm31_constant(cs, v):                                                # qm31.rs:35-73
    v = tuple(x % P for x in v)
    if v == Q0:
        return qm31_zero(cs)
    if v == Q1:
        return qm31_one(cs)
    if v == (0, 1, 0, 0):
        return V(cs, v, 2, 2)
    if v == (0, 0, 1, 0):
        return V(cs, v, 3, 2)
    key = "qm31 %d,%d,%d,%d" % v
    if key in cs.cache:
        return V(cs, v, cs.cache[key], 2)
    var = cs.new_qm31(v, "constant")
    cs.cache[key] = var
    return V(cs, v, var, 2)


def qm31_from_m31(a0, a1, a2, a3):                                       # qm31.rs:245-256
    cs = a0.cs
    l = cs.add(a0.variable, cs.mul(a1.variable, 2))
    r = cs.mul(cs.add(a2.variable, cs.mul(a3.variable, 2)), 3)
    return V(cs, (a0.value[0], a1.value[0], a2.value[0], a3.value[0]), cs.add(l, r), 2)


def qm31_decompose_m31(x):                                               # qm31.rs:258-272
    cs = x.cs
    a = [m31_witness(cs, x.value[k], (LOG_COORD, x.variable, k)) for k in range(4)]
    l = cs.add(a[0].variable, cs.mul(a[1].variable, 2))
    r = cs.mul(cs.add(a[2].variable, cs.mul(a[3].variable, 2)), 3)
    cs.insert_gate(l, r, x.variable, 1)
    return a


def qm31_decompose_cm31(x):                                              # qm31.rs:274-281
    v = qm31_decompose_m31(x)
    a0 = v[1].shift_by_i() + v[0]                                        # CM31Var::from(&v[1]).shift_by_i() + &v[0]
    a1 = v[3].shift_by_i() + v[2]
    return a0, a1


def qm31_inv(x):                                                         # qm31.rs:352-359
    r = qm31_witness(x.cs, q_inv(x.value), (LOG_INV_QM31, x.variable, 0))
    x.cs.insert_gate(x.variable, r.variable, 1, 0)
    return r


def qm31_swap(a, b, bit_value, bit_variable):                            # qm31.rs:437-464
    cs = a.cs
    lv, rv = (b.value, a.value) if bit_value else (a.value, b.value)
    b_minus_a = b - a
    left = cs.mul(b_minus_a.variable, bit_variable)
    right = cs.mul_constant(left, P - 1)
    left = cs.add(a.variable, left)
    right = cs.add(b.variable, right)
    return V(cs, lv, left, 2), V(cs, rv, right, 2)


# ---- bits (primitives/bits/src/lib.rs) -------------------------------------------------------------------------------
class Bits:
    def __init__(self, cs, value, variables):
        self.cs, self.value, self.variables = cs, list(value), list(variables)

    @staticmethod
    def new_witness(cs, bools, of=None):                                 # :25-43
        variables = []
        for k, b in enumerate(bools):
            bit = cs.new_qm31(Q1 if b else Q0, "witness", None if of is None else (LOG_BIT, of, k))
            variables.append(bit)
            minus_one = m31_constant(cs, P - 1)
            bm1 = cs.add(bit, minus_one.variable)
            cs.insert_gate(bit, bm1, 0, 0)
        return Bits(cs, bools, variables)

    @staticmethod
    def from_m31(v, l):                                                  # :48-82
        cs = v.cs
        cur = v.value[0]
        bools = [(cur >> k) & 1 != 0 for k in range(l)]
        res = Bits.new_witness(cs, bools, v.variable)
        rec = V(cs, qm(1 if res.value[0] else 0), res.variables[0], 0)
        for i in range(1, l):
            t = V(cs, qm(1 if res.value[i] else 0), res.variables[i], 0).mul_constant_m31(1 << i)
            rec = rec + t
        rec.equalverify(v)
        if l == 31:
            prod = cs.mul(res.variables[0], res.variables[1])
            for i in range(2, l):
                prod = cs.mul(prod, res.variables[i])
            cs.enforce_zero(prod)
        return res

    def get_value(self):
        return sum(1 << k for k, b in enumerate(self.value) if b)

    def compose_range(self, lo, hi):                                     # :96-118
        cs = self.cs
        s = 1 if self.value[lo] else 0
        var = self.variables[lo]
        for shift, i in enumerate(range(lo + 1, hi)):
            if self.value[i]:
                s += 1 << (shift + 1)
            sv = cs.mul_constant(self.variables[i], 1 << (shift + 1))
            var = cs.add(var, sv)
        return V(cs, qm(s), var, 0)

    def index_range(self, lo, hi=None):
        return Bits(self.cs, self.value[lo:hi], self.variables[lo:hi])


# ---- Poseidon2 half states (primitives/poseidon31/src/lib.rs, native variant) ----------------------------------------
class Half:
    __slots__ = ("cs", "value", "left_variable", "right_variable", "sel_value")

    def __init__(self, cs, value, l, r, sel):
        self.cs, self.value, self.left_variable, self.right_variable, self.sel_value = cs, list(value), l, r, sel

    @staticmethod
    def single_use_witness_only(cs, value):                              # :51-74
        if cs.kind == "without":
            return HalfE(cs, [qm31_witness(cs, value[0:4]), qm31_witness(cs, value[4:8])])
        return Half(cs, value, 0, 0, 0)

    @staticmethod
    def from_m31(s):                                                     # :76-105
        cs = s[0].cs
        left = qm31_from_m31(s[0], s[1], s[2], s[3])
        right = qm31_from_m31(s[4], s[5], s[6], s[7])
        if cs.kind == "without":
            return HalfE(cs, [left, right])
        sel = cs.assemble_poseidon_gate(left.variable, right.variable)
        return Half(cs, [x.value[0] for x in s], left.variable, right.variable, sel)

    @staticmethod
    def from_qm31(a, b):                                                 # :107-131
        cs = a.cs
        if cs.kind == "without":
            return HalfE(cs, [a, b])
        sel = cs.assemble_poseidon_gate(a.variable, b.variable)
        return Half(cs, list(a.value) + list(b.value), a.variable, b.variable, sel)

    @staticmethod
    def new_variables(cs, value, mode):                                  # :142-188
        mk = {"witness": qm31_witness, "input": qm31_input}[mode]
        left, right = mk(cs, value[0:4]), mk(cs, value[4:8])
        if cs.kind == "without":
            return HalfE(cs, [left, right])
        return Half(cs, value, left.variable, right.variable, cs.assemble_poseidon_gate(left.variable, right.variable))

    @staticmethod
    def new_witness(cs, value, src=None):                                # :146-166 (QM31 witnesses cost no rows)
        left = qm31_witness(cs, value[0:4], src)
        right = qm31_witness(cs, value[4:8], src)
        if cs.kind == "without":
            return HalfE(cs, [left, right])
        sel = cs.assemble_poseidon_gate(left.variable, right.variable)
        return Half(cs, value, left.variable, right.variable, sel)

    @staticmethod
    def zero(cs):                                                        # :191-218
        if cs.kind == "without":
            return HalfE(cs, [qm31_zero(cs), qm31_zero(cs)])
        if "poseidon2 zero_half" not in cs.cache:
            cs.cache["poseidon2 zero_half"] = cs.assemble_poseidon_gate(0, 0)
        return Half(cs, [0] * 8, 0, 0, cs.cache["poseidon2 zero_half"])

    def to_qm31(self):                                                   # :220-249
        return [V(self.cs, tuple(self.value[0:4]), self.left_variable, 2), V(self.cs, tuple(self.value[4:8]), self.right_variable, 2)]

    def equalverify(self, o):                                            # :425-438
        self.cs.insert_gate(self.left_variable, 0, o.left_variable, 1)
        self.cs.insert_gate(self.right_variable, 0, o.right_variable, 1)


class HalfE:
    """Poseidon2HalfEmulatedVar (poseidon31/src/lib.rs:30-34): two QM31 limbs"""
    __slots__ = ("cs", "elems")

    def __init__(self, cs, elems):
        self.cs, self.elems = cs, list(elems)

    @property
    def value(self):
        return list(self.elems[0].value) + list(self.elems[1].value)

    def to_qm31(self):
        return list(self.elems)

    def equalverify(self, o):
        for a, b in zip(self.elems, o.elems):
            a.equalverify(b)


def qm31_input(cs, v):
    return V(cs, tuple(v), cs.new_qm31(v, "input"), 2)


def m31_input(cs, v):
    return V(cs, qm(v), cs.new_m31(v % P, "input"), 0)


def _e_m4(x):                                                            # emulated.rs:12-22
    cs = x.cs
    k = qm31_constant(cs, (1, 1, 1, 1))
    var = cs.do_m4_gate(x.variable, k.variable)
    return V(cs, cs.variables[var], var, 2)


def _e_mds16(st):                                                        # emulated.rs:24-35
    p = [_e_m4(x) for x in st]
    t = p[0] + p[1]
    t = t + p[2]
    t = t + p[3]
    return [p[0] + t, p[1] + t, p[2] + t, p[3] + t]


def _pow4_witness(cs, var):
    b = tuple(pow(x, 4, P) for x in cs.variables[var])
    return qm31_witness(cs, b, (LOG_POW4, var, 0))


def _e_pow5m4(x):                                                        # emulated.rs:37-61
    cs = x.cs
    b = _pow4_witness(cs, x.variable)
    var = cs.do_pow5m4_gate(x.variable, b.variable)
    return V(cs, cs.variables[var], var, 2)


def _e_pow5(cs, var):                                                    # emulated.rs:63-78
    b = _pow4_witness(cs, var)
    return cs.do_pow5_gate(var, b.variable)


def poseidon_permute_emulated(left, right, is_swap=None):                # emulated.rs:80-221
    cs = left.cs
    K = p2_constants()
    if is_swap is not None:
        bit = V(cs, qm(1 if is_swap[0] else 0), is_swap[1], 0)
        rml = [right.elems[i] - left.elems[i] for i in range(2)]
        rmlb = [rml[i] * bit for i in range(2)]
        nl = [rmlb[i] + left.elems[i] for i in range(2)]
        nr = [right.elems[i] - rmlb[i] for i in range(2)]
    else:
        nl, nr = list(left.elems), list(right.elems)
    st = [nl[0], nl[1], nr[0], nr[1]]
    st = _e_mds16(st)

    def full_rounds(st, rc):
        for r in range(4):
            for i in range(4):
                st[i] = st[i] + qm31_constant(cs, tuple(rc[16 * r + 4 * i:16 * r + 4 * i + 4]))
            for i in range(4):
                st[i] = _e_pow5m4(st[i])
            t = st[0] + st[1]
            t = t + st[2]
            t = t + st[3]
            st = [st[0] + t, st[1] + t, st[2] + t, st[3] + t]
        return st
    st = full_rounds(st, K["RC_FIRST"])
    for r in range(14):
        first_only = cs.do_hadamard(st[0].variable, 1)
        k = qm31_constant(cs, (0, 1, 1, 1))
        without_first = cs.do_hadamard(st[0].variable, k.variable)
        k = m31_constant(cs, K["RC_PARTIAL"][r])
        first_only = cs.add(first_only, k.variable)
        first_only = _e_pow5(cs, first_only)
        tmp = cs.add(first_only, without_first)
        st[0] = V(cs, cs.variables[tmp], tmp, 2)
        s1 = cs.do_grandsum_gate(st[0].variable, st[1].variable)
        s2 = cs.do_grandsum_gate(st[2].variable, st[3].variable)
        total = cs.add(s1, s2)
        for i in range(4):
            k = qm31_constant(cs, tuple(K["DIAG16"][4 * i:4 * i + 4]))
            v = cs.do_hadamard(st[i].variable, k.variable)
            v = cs.add(total, v)
            st[i] = V(cs, cs.variables[v], v, 2)
    st = full_rounds(st, K["RC_LAST"])
    cs.n_perm += 1
    return HalfE(cs, [st[0], st[1]]), HalfE(cs, [st[2], st[3]])


def permute(left, right, ignore_left, ignore_right, is_swap=None):       # poseidon31/src/lib.rs:282-407
    if isinstance(left, HalfE):
        return poseidon_permute_emulated(left, right, is_swap)
    cs = left.cs
    swap = is_swap is not None and is_swap[0]
    state = (right.value + left.value) if swap else (left.value + right.value)
    state = poseidon2_permute(state)
    cs.n_perm += 1
    # value log: a half is two variables, or a literal (a sibling hash that exists only as a value)
    rec = []
    for h in (left, right):
        if h.sel_value == 0 and (h.left_variable, h.right_variable) == (0, 0) and any(h.value):
            rec += [1, 0, 0]
        else:
            rec += [0, h.left_variable, h.right_variable]
    rec.append(is_swap[1] if is_swap is not None else NO_VAR)
    cs.log.append((LOG_PERM, len(cs.perm_log), 0, 0, 0, 0, 0, 0))
    outs, out_vars = [], []
    for ignore, half in ((ignore_left, state[0:8]), (ignore_right, state[8:16])):
        if ignore:
            outs.append(Half(cs, half, 0, 0, 0))
            out_vars += [NO_VAR, NO_VAR]
        else:
            outs.append(Half.new_witness(cs, half, "perm"))
            out_vars += [outs[-1].left_variable, outs[-1].right_variable]
    cs.perm_log.append(rec + out_vars + [0] * 5 + list(left.value) + list(right.value))
    e = [(h.sel_value, list(h.value)) for h in (left, right, outs[0], outs[1])]
    cs.invoke_poseidon_accelerator(e[0], e[1], e[2], e[3], (is_swap[1], bool(is_swap[0])) if is_swap is not None else (0, False))
    return outs[0], outs[1]


def permute_get_rate(l, r, is_swap=None):
    return permute(l, r, False, True, is_swap)[0]


def permute_get_capacity(l, r, is_swap=None):
    return permute(l, r, True, False, is_swap)[1]


# ---- Merkle hasher (primitives/merkle/src/lib.rs) --------------------------------------------------------------------
def _sponge(items, width, mk):
    """shared chunk walk of hash_m31_columns_get_capacity (:140-180) / hash_qm31_columns_get_capacity (:98-138)"""
    cs = items[0].cs
    n = len(items)
    num_chunk = (n + width - 1) // width
    pad = [m31_zero(cs) for _ in range(width)]
    first = list(items[0:min(n, width)]) + pad[min(n, width):]
    z = Half.zero(cs)
    first_chunk = mk(first)
    digest = permute_get_capacity(first_chunk, z)
    if num_chunk == 1:
        return digest
    for c in range(1, num_chunk - 1):
        digest = permute_get_capacity(mk(list(items[c * width:(c + 1) * width])), digest)
    remain = n % width
    last = list(items[n - width:]) if remain == 0 else list(items[n - remain:]) + pad[remain:]
    return permute_get_capacity(mk(last), digest)


def hash_m31_columns_get_capacity(m31):
    return _sponge(m31, 8, Half.from_m31)


def hash_m31_columns_get_rate(m31):                                      # :51-90
    digest = _sponge(m31, 8, Half.from_m31)
    return permute_get_rate(Half.zero(m31[0].cs), digest)


def hash_qm31_columns_get_capacity(q):
    return _sponge(q, 2, lambda s: Half.from_qm31(s[0], s[1]))


def hash_qm31_columns_get_rate(q):                                       # :92-96
    digest = hash_qm31_columns_get_capacity(q)
    return permute_get_rate(Half.zero(q[0].cs), digest)


# ---- channel (primitives/channel/src/lib.rs) -------------------------------------------------------------------------
class Channel:
    def __init__(self, cs):
        self.cs, self.n_sent, self.digest = cs, 0, Half.zero(cs)

    def mix_root(self, root):
        self.digest = permute_get_capacity(root, self.digest)
        self.n_sent = 0

    def draw_felts(self):
        n_sent = m31_constant(self.cs, self.n_sent)
        self.n_sent += 1
        left = Half.from_qm31(n_sent, qm31_zero(self.cs))
        return permute_get_rate(left, self.digest).to_qm31()

    def mix_one_felt(self, f):
        self.digest = permute_get_capacity(Half.from_qm31(f, qm31_zero(self.cs)), self.digest)
        self.n_sent = 0

    def mix_two_felts(self, a, b):
        self.digest = permute_get_capacity(Half.from_qm31(a, b), self.digest)
        self.n_sent = 0


# ---- circle points (primitives/circle/src/lib.rs) --------------------------------------------------------------------
class PointM31:
    def __init__(self, x, y):
        self.x, self.y = x, y

    def __add__(self, o):                                                # :46-57
        x1x2 = self.x * o.x
        y1y2 = self.y * o.y
        x1y2 = self.x * o.y
        y1x2 = self.y * o.x
        return PointM31(x1x2 - y1y2, x1y2 + y1x2)

    def double(self):                                                    # :61-69
        xx = self.x * self.x
        yy = self.y * self.y
        xy = self.x * self.y
        return PointM31(xx - yy, xy.mul_constant_m31(2))

    @staticmethod
    def new_constant(cs, p):
        return PointM31(m31_constant(cs, p[0]), m31_constant(cs, p[1]))

    @staticmethod
    def select(cs, point, bit_value, bit_variable):                      # :74-104 (the op constant follows the SELECTED value)
        value = point if bit_value else (1, 0)
        nx = cs.mul_constant(bit_variable, (value[0] - 1) % P)
        nx = cs.add(nx, 1)
        ny = cs.mul_constant(bit_variable, value[1])
        return PointM31(V(cs, qm(value[0]), nx, 0), V(cs, qm(value[1]), ny, 0))

    def conditional_negate(self, bit_value, bit_variable):               # :106-130
        cs = self.x.cs
        yv = (P - self.y.value[0]) % P if bit_value else self.y.value[0]
        m = cs.mul_constant(bit_variable, P - 2)
        m = cs.add(m, 1)
        return PointM31(self.x, V(cs, qm(yv), cs.mul(m, self.y.variable), 0))


class PointQM31:
    def __init__(self, x, y):
        self.x, self.y = x, y

    @staticmethod
    def from_t(t):                                                       # :204-219
        cs = t.cs
        t_doubled = t + t
        t_squared = t * t
        tp1 = t_squared + m31_one(cs)
        tp1_inv = qm31_inv(tp1)
        omt = (-t_squared) + m31_one(cs)
        return PointQM31(omt * tp1_inv, t_doubled * tp1_inv)

    def repeated_double_x_only(self, n):                                 # :226-234
        x = self.x
        for _ in range(n):
            sq = x * x
            x = (sq + sq) - m31_one(x.cs)
        return x

    def add_m31_point(self, p):                                          # :236-250
        x1x2 = self.x.mul_constant_m31(p[0])
        y1y2 = self.y.mul_constant_m31(p[1])
        x1y2 = self.x.mul_constant_m31(p[1])
        y1x2 = self.y.mul_constant_m31(p[0])
        return PointQM31(x1x2 - y1y2, x1y2 + y1x2)


# ---- queries (primitives/query/src/lib.rs) ---------------------------------------------------------------------------
class PointCarryingQuery:
    def __init__(self, bits, last_step, point):
        self.bits, self.last_step, self.point = bits, last_step, point

    @staticmethod
    def new(bits):                                                       # :56-139
        cs = bits.cs
        log_size = len(bits.value)
        initial, step = cp_gen(log_size + 2), cp_gen(log_size)          # CanonicCoset(log+1).circle_domain().half_coset
        steps = []
        cur = step
        for _ in range(log_size - 1):
            steps.append(cur)
            cur = cp_double(cur)
        combs = list(zip(steps, reversed(bits.value[1:]), reversed(bits.variables[1:])))
        cur = PointM31.new_constant(cs, initial)
        for k in range(0, len(combs), 2):
            chunk = combs[k:k + 2]
            if len(chunk) == 1:
                point = PointM31.select(cs, chunk[0][0], chunk[0][1], chunk[0][2])
                cur = point + cur
            else:
                p00, p01, p10 = (1, 0), chunk[0][0], chunk[1][0]
                p11 = cp_add(p01, p10)
                value = {(False, False): p00, (True, False): p01, (False, True): p10, (True, True): p11}[(chunk[0][1], chunk[1][1])]
                a, b = chunk[0][2], chunk[1][2]
                oma = cs.add(1, cs.mul_constant(a, P - 1))
                omb = cs.add(1, cs.mul_constant(b, P - 1))
                b00 = cs.mul(oma, omb)
                b01 = cs.mul(a, omb)
                b10 = cs.mul(oma, b)
                b11 = cs.mul(a, b)
                x = cs.mul_constant(b00, p00[0])
                x = cs.add(x, cs.mul_constant(b01, p01[0]))
                x = cs.add(x, cs.mul_constant(b10, p10[0]))
                x = cs.add(x, cs.mul_constant(b11, p11[0]))
                y = cs.mul_constant(b00, p00[1])
                y = cs.add(y, cs.mul_constant(b01, p01[1]))
                y = cs.add(y, cs.mul_constant(b10, p10[1]))
                y = cs.add(y, cs.mul_constant(b11, p11[1]))
                point = PointM31(V(cs, qm(value[0]), x, 0), V(cs, qm(value[1]), y, 0))
                cur = point + cur
        return PointCarryingQuery(bits, cp_neg(steps[-1]), cur)

    def get_next_point(self):                                            # :140-144
        return self.point.double().conditional_negate(self.bits.value[0], self.bits.variables[0])

    def get_next_point_x(self):                                          # :145-149
        xx = self.point.x * self.point.x
        yy = self.point.y * self.point.y
        return xx - yy

    def next(self):                                                      # :150-162
        cs = self.bits.cs
        t = PointM31.select(cs, self.last_step, self.bits.value[1], self.bits.variables[1])
        return PointCarryingQuery(self.bits.index_range(1), self.last_step, (self.point + t).double())


def query_positions_per_log_size(lo, hi, raw_queries):                   # :19-38
    elems = [PointCarryingQuery.new(Bits.from_m31(r, 31).index_range(0, hi)) for r in raw_queries]
    points = {hi: elems}
    for log_size in range(hi - 1, lo - 1, -1):
        elems = [e.next() for e in elems]
        points[log_size] = elems
    return points


# ---- line polynomial (primitives/line/src/lib.rs:39-67) --------------------------------------------------------------
def line_poly_eval_at_point(cs, coeffs, x):
    lg = len(coeffs).bit_length() - 1
    doublings = [x]
    for _ in range(1, lg):
        x_sq = x * x
        x = x_sq + x_sq
        x = x + m31_constant(cs, P - 1)
        doublings.append(x)

    def fold(values, factors):
        n = len(values)
        if n == 1:
            return values[0]
        l = fold(values[:n // 2], factors[1:])
        r = fold(values[n // 2:], factors[1:])
        return l + (r * factors[0])
    return fold(coeffs, doublings)


# ---- proof wire format -------------------------------------------------------------------------------
class Proof:
    pass


def parse_proof(blob):
    w = np.frombuffer(blob, dtype="<u4")
    pos = [0]

    def u32():
        v = int(w[pos[0]]); pos[0] += 1
        return v

    def u64():
        v = int(w[pos[0]]) | (int(w[pos[0] + 1]) << 32); pos[0] += 2
        return v

    def words(n):
        v = [int(x) for x in w[pos[0]:pos[0] + n]]; pos[0] += n
        return v

    def decommitment():
        n = u64(); hw = [words(8) for _ in range(n)]
        n = u64(); cw = words(n)
        return hw, cw

    def layer():
        n = u64(); fw = [tuple(words(4)) for _ in range(n)]
        d = decommitment()
        return {"fri_witness": fw, "decommitment": d, "commitment": words(8)}
    p = Proof()
    p.log_size_plonk, p.log_size_poseidon = u32(), u32()
    p.plonk_total_sum, p.poseidon_total_sum = tuple(words(4)), tuple(words(4))
    p.pow_bits, p.log_blowup, p.log_last = u32(), u32(), u32()
    p.n_queries = u64()
    p.commitments = [words(8) for _ in range(u64())]
    p.sampled_values = [[[tuple(words(4)) for _ in range(u64())] for _ in range(u64())] for _ in range(u64())]
    p.decommitments = [decommitment() for _ in range(u64())]
    p.queried_values = [words(u64()) for _ in range(u64())]
    p.proof_of_work = u64()
    p.first_layer = layer()
    p.inner_layers = [layer() for _ in range(u64())]
    p.last_coeffs = [tuple(words(4)) for _ in range(u64())]
    p.last_log_size = u32()
    assert pos[0] == len(w), "trailing bytes"
    return p


# ---- hints from the C oracle -----------------------------------------------------------------------------------------
MAX_INNER, MAX_Q = 32, 128
_Q4 = ctypes.c_uint32 * 4


class Hints(ctypes.Structure):
    """orc_hints (oracle/orc.h)"""
    _fields_ = [("single_depth", ctypes.c_uint32 * 4), ("single_ncols", (ctypes.c_uint32 * 33) * 4),
                ("single_cols", ((ctypes.c_uint32 * 64) * MAX_Q) * 4), ("single_sib", (((ctypes.c_uint32 * 8) * 32) * MAX_Q) * 4),
                ("pair_depth", ctypes.c_uint32 * (1 + MAX_INNER)), ("pair_has_data", (ctypes.c_uint8 * 33) * (1 + MAX_INNER)),
                ("pair_self", ((_Q4 * 33) * MAX_Q) * (1 + MAX_INNER)), ("pair_sib", ((_Q4 * 33) * MAX_Q) * (1 + MAX_INNER)),
                ("pair_sib_hash", (((ctypes.c_uint32 * 8) * 32) * MAX_Q) * (1 + MAX_INNER))]


class SinglePath:
    """SinglePathMerkleProof (components/hints/src/decommit.rs:10-16)"""
    def __init__(self, h, t, i, query):
        self.depth = int(h.single_depth[t])
        self.query = query
        self.sibling_hashes = [list(h.single_sib[t][i][k]) for k in range(self.depth)]
        self.columns = {}
        off = 0
        for layer in range(self.depth, -1, -1):
            n = int(h.single_ncols[t][layer])
            if n:
                self.columns[layer] = [int(x) for x in h.single_cols[t][i][off:off + n]]
                off += n


class SinglePair:
    """SinglePairMerkleProof (components/hints/src/folding.rs:21-28)"""
    def __init__(self, h, l, i, query):
        self.depth = int(h.pair_depth[l])
        self.query = query
        self.sibling_hashes = [list(h.pair_sib_hash[l][i][k]) for k in range(self.depth - 1)]
        self.self_columns, self.siblings_columns = {}, {}
        for layer in range(self.depth + 1):
            if h.pair_has_data[l][layer]:
                self.self_columns[layer] = tuple(h.pair_self[l][i][layer])
                self.siblings_columns[layer] = tuple(h.pair_sib[l][i][layer])


def compute_hints(blob, inputs, verify_out_cls):
    """runs the C oracle's native verifier; returns (VerifyOut, Hints)"""
    lib = _orc()
    buf = np.zeros((len(blob) + 3) // 4 * 4, dtype=np.uint8)
    buf[:len(blob)] = np.frombuffer(blob, dtype=np.uint8)
    idx = np.array(inputs[0], dtype=np.uint32)
    vals = np.array(inputs[1], dtype=np.uint32)
    out, h = verify_out_cls(), Hints()
    lib.orc_verify_proof_hints(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(blob)), idx.ctypes.data_as(ctypes.c_void_p),
                               vals.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(idx.size), ctypes.byref(out), ctypes.byref(h))
    return out, h


# ---- the verifier circuit --------------------------------------------------------------------------------------------
class ProofVar:
    """PlonkWithPoseidonProofVar::new_witness (components/recursive/data_structures/src/lib.rs:30-213)"""
    def __init__(self, cs, p):
        self.cs, self.p = cs, p
        self.log_size_plonk = m31_witness(cs, p.log_size_plonk)
        self.log_size_poseidon = m31_witness(cs, p.log_size_poseidon)
        self.plonk_total_sum = qm31_witness(cs, p.plonk_total_sum)
        self.poseidon_total_sum = qm31_witness(cs, p.poseidon_total_sum)
        self.commitments = [Half.new_witness(cs, c) for c in p.commitments]
        self.sampled_values = [[[qm31_witness(cs, v) for v in col] for col in tree] for tree in p.sampled_values]
        self.first_layer_commitment = Half.new_witness(cs, p.first_layer["commitment"])
        self.inner_layer_commitments = [Half.new_witness(cs, l["commitment"]) for l in p.inner_layers]
        self.last_poly = [qm31_witness(cs, c) for c in p.last_coeffs]
        n = p.proof_of_work
        self.proof_of_work = [m31_witness(cs, n & ((1 << 22) - 1)), m31_witness(cs, (n >> 22) & ((1 << 21) - 1)),
                              m31_witness(cs, (n >> 43) & ((1 << 21) - 1))]


class Shape:
    """value-independent facts FiatShamirHints carries (components/hints/src/fiat_shamir.rs:93-104,239-305)"""
    def __init__(self, p):
        self.log_plonk, self.log_poseidon, self.blowup = p.log_size_plonk, p.log_size_poseidon, p.log_blowup
        self.n_queries, self.pow_bits, self.log_last = p.n_queries, p.pow_bits, p.log_last
        self.max_first = p.log_last + p.log_blowup + 1 + len(p.inner_layers)
        self.all_log_sizes = sorted({self.log_plonk + self.blowup, self.log_poseidon + self.blowup, self.max_first})
        self.composition_log_degree_bound = self.max_first - self.blowup + 1
        # column -> (blown-up log size, component log size); tree 0..2 split Plonk | Poseidon, tree 3 = composition
        split = [10, 12, 8]
        self.column_log_sizes, self.column_component = [], []
        for t in range(4):
            n = len(p.sampled_values[t])
            if t == 3:
                self.column_log_sizes.append([self.max_first] * n)
                self.column_component.append([None] * n)
            else:
                self.column_log_sizes.append([self.log_plonk + self.blowup] * split[t] + [self.log_poseidon + self.blowup] * (n - split[t]))
                self.column_component.append(["plonk"] * split[t] + ["poseidon"] * (n - split[t]))
        self.split = split


def fiat_shamir(cs, pv, shape, inputs):
    """FiatShamirResults::compute (components/recursive/fiat_shamir/src/lib.rs:31-176)"""
    ch = Channel(cs)
    ch.mix_root(pv.commitments[0])
    ch.mix_one_felt(pv.log_size_plonk)
    ch.mix_one_felt(pv.log_size_poseidon)
    ch.mix_root(pv.commitments[1])
    z, alpha = ch.draw_felts()
    alpha_powers = [qm31_one(cs), alpha, alpha * alpha]                       # LookupElementsVar::from_z_and_alpha
    ch.mix_two_felts(pv.plonk_total_sum, pv.poseidon_total_sum)
    ch.mix_root(pv.commitments[2])
    random_coeff = ch.draw_felts()[0]
    ch.mix_root(pv.commitments[3])
    oods_point = PointQM31.from_t(ch.draw_felts()[0])
    flat = [v for tree in pv.sampled_values for col in tree for v in col]
    for k in range(0, len(flat), 2):
        if k + 1 == len(flat):
            ch.mix_one_felt(flat[k])
        else:
            ch.mix_two_felts(flat[k], flat[k + 1])
    after = ch.draw_felts()[0]
    fri_alphas = []
    ch.mix_root(pv.first_layer_commitment)
    fri_alphas.append(ch.draw_felts()[0])
    for l in pv.inner_layer_commitments:
        ch.mix_root(l)
        fri_alphas.append(ch.draw_felts()[0])
    for k in range(0, len(pv.last_poly), 2):
        if k + 1 == len(pv.last_poly):
            ch.mix_one_felt(pv.last_poly[k])
        else:
            ch.mix_two_felts(pv.last_poly[k], pv.last_poly[k + 1])
    nonce_felt = qm31_from_m31(pv.proof_of_work[0], pv.proof_of_work[1], pv.proof_of_work[2], m31_zero(cs))
    Bits.from_m31(pv.proof_of_work[0], 22)
    Bits.from_m31(pv.proof_of_work[1], 21)
    Bits.from_m31(pv.proof_of_work[2], 21)
    ch.mix_one_felt(nonce_felt)
    lower = Bits.from_m31(qm31_decompose_m31(ch.digest.to_qm31()[0])[0], 31).compose_range(0, shape.pow_bits)
    lower.equalverify(m31_zero(cs))
    felts = []
    for _ in range((shape.n_queries + 3) // 4):
        a, b = ch.draw_felts()
        felts += [a, b]
    raw_queries = []
    for f in felts:
        raw_queries += qm31_decompose_m31(f)
    raw_queries = raw_queries[:shape.n_queries]
    input_sum = qm31_zero(cs)
    for idx, v in inputs:
        s = (v + (qm31_constant(cs, qm(idx)) * alpha)) - z
        input_sum = input_sum + qm31_inv(s)
    ((input_sum + pv.poseidon_total_sum) + pv.plonk_total_sum).equalverify(qm31_zero(cs))
    return dict(z=z, alpha=alpha, alpha_powers=alpha_powers, random_coeff=random_coeff, after=after, oods_point=oods_point,
                raw_queries=raw_queries, fri_alphas=fri_alphas)


def coset_vanishing(pt, log_size):                                       # composition/src/lib.rs:18-29
    cs = pt.x.cs
    x = pt.add_m31_point((1, 0)).x                                       # -initial + step/2 is the identity for canonic cosets
    for _ in range(1, log_size):
        sq = x * x
        x = (sq + sq) - m31_one(cs)
    return x


class EvalAtRow:
    """composition/src/data_structures.rs:82-215"""
    def __init__(self, mask, total_sum, denom_inverse, log_size, acc):
        self.col_index = [0, 0, 0, 0]
        self.mask, self.denom_inverse, self.acc = mask, denom_inverse, acc
        self.cumsum_shift = total_sum.mul_constant_m31(m_inv(1 << log_size))
        self.fracs = []

    def next_mask(self, interaction):
        k = self.col_index[interaction]
        self.col_index[interaction] += 1
        return self.mask[interaction][k]

    def next_trace_mask(self):
        return self.next_mask(1)[0]

    def preprocessed(self):
        return self.next_mask(0)[0]

    def next_extension_mask(self, interaction, n):
        cols = [self.next_mask(interaction) for _ in range(4)]
        assert all(len(c) == n for c in cols)
        return [combine_ef([c[k] for c in cols]) for k in range(n)]

    def add_to_relation(self, rel, multiplicity, values):               # :148-165
        denom = rel["alpha_powers"][0] * values[0]
        for ap, v in list(zip(rel["alpha_powers"], values))[1:]:
            denom = denom + (ap * v)
        denom = denom - rel["z"]
        self.fracs.append((multiplicity, denom))

    def add_constraint(self, value):                                     # :167-170, :25-27
        ev = value * self.denom_inverse
        self.acc[0] = (self.acc[0] * self.acc[1]) + ev

    def finalize_logup(self, batch):                                     # :172-210
        cs = self.denom_inverse.cs
        n_batches = (len(self.fracs) + batch - 1) // batch
        batched = []
        for k in range(0, len(self.fracs), batch):
            chunk = self.fracs[k:k + batch]
            if len(chunk) == 1:
                batched.append(chunk[0])
            else:
                p, q = chunk[0]
                for e in chunk[1:]:
                    p = (p * e[1]) + (e[0] * q)
                    q = q * e[1]
                batched.append((p, q))
        prev_col = qm31_zero(cs)
        for num, den in batched[:n_batches - 1]:
            cur = self.next_extension_mask(2, 1)[0]
            diff = cur - prev_col
            prev_col = cur
            self.add_constraint((diff * den) - num)
        for num, den in batched[n_batches - 1:]:
            prev_row, cur = self.next_extension_mask(2, 2)
            diff = (cur - prev_row) - prev_col
            fixed = diff + self.cumsum_shift
            self.add_constraint((fixed * den) - num)


def combine_ef(v):                                                       # data_structures.rs:143-146
    return ((v[0] + v[1].shift_by_i()) + v[2].shift_by_j()) + v[3].shift_by_ij()


def evaluate_plonk(cs, rel, ev):                                         # composition/src/plonk.rs:8-82
    a_wire, b_wire, c_wire, op = ev.preprocessed(), ev.preprocessed(), ev.preprocessed(), ev.preprocessed()
    mult_a, mult_b, mult_c = ev.preprocessed(), ev.preprocessed(), ev.preprocessed()
    poseidon_wire, mult_poseidon, enforce_c_m31 = ev.preprocessed(), ev.preprocessed(), ev.preprocessed()
    av = [ev.next_trace_mask() for _ in range(4)]
    bv = [ev.next_trace_mask() for _ in range(4)]
    cv = [ev.next_trace_mask() for _ in range(4)]
    ev.add_constraint(enforce_c_m31 * cv[1])
    ev.add_constraint(enforce_c_m31 * cv[2])
    ev.add_constraint(enforce_c_m31 * cv[3])
    a_val, b_val, c_val = combine_ef(av), combine_ef(bv), combine_ef(cv)
    ev.add_constraint((c_val - (op * (a_val + b_val))) - (((qm31_one(cs) - op) * a_val) * b_val))
    ev.add_to_relation(rel, mult_a, [a_val, a_wire])
    ev.add_to_relation(rel, mult_b, [b_val, b_wire])
    ev.add_to_relation(rel, mult_c, [c_val, c_wire])
    ev.add_to_relation(rel, -mult_poseidon, [poseidon_wire, a_val, b_val])
    ev.finalize_logup(2)


def _apply_m4(x):                                                        # composition/src/poseidon.rs:10-24
    t0 = x[0] + x[1]
    t02 = t0 + t0
    t1 = x[2] + x[3]
    t12 = t1 + t1
    t2 = (x[1] + x[1]) + t1
    t3 = (x[3] + x[3]) + t0
    t4 = (t12 + t12) + t3
    t5 = (t02 + t02) + t2
    t6 = t3 + t5
    t7 = t2 + t4
    return [t6, t5, t7, t4]


def _external(s):                                                        # :28-50
    for i in range(4):
        s[4 * i:4 * i + 4] = _apply_m4(s[4 * i:4 * i + 4])
    for j in range(4):
        t = ((s[j] + s[j + 4]) + s[j + 8]) + s[j + 12]
        for i in range(4):
            s[4 * i + j] = s[4 * i + j] + t


def _internal(s):                                                        # :55-66
    total = s[0]
    for x in s[1:]:
        total = total + x
    s[0] = s[0] + ((s[0] + s[0]) + total)
    for i in range(1, 16):
        s[i] = s[i].mul_constant_m31(1 << (i + 1)) + total


def _pow5(x):
    x2 = x * x
    x4 = x2 * x2
    return x4 * x


def evaluate_poseidon(cs, rel, ev):                                      # composition/src/poseidon.rs:73-241
    is_first, is_last, is_full = ev.preprocessed(), ev.preprocessed(), ev.preprocessed()
    o = qm31_one(cs)
    not_first = o - is_first
    not_last = o - is_last
    is_partial = not_first - is_full
    round_id = ev.preprocessed()
    rc0 = [ev.preprocessed() for _ in range(16)]
    rc1 = [ev.preprocessed() for _ in range(16)]
    ext1, ext2, ext1_nz, ext2_nz = ev.preprocessed(), ev.preprocessed(), ev.preprocessed(), ev.preprocessed()
    swap_bit_addr = rc0[0]
    in_state = [ev.next_trace_mask() for _ in range(16)]
    mid = [ev.next_trace_mask() for _ in range(16)]
    out_state = [ev.next_trace_mask() for _ in range(16)]
    swap_bit = mid[0]
    om_swap = o - swap_bit
    perm = []
    for i in range(16):
        if i < 8:
            perm.append((in_state[i] * om_swap) + (in_state[i + 8] * swap_bit))
        else:
            perm.append((in_state[i - 8] * swap_bit) + (in_state[i] * om_swap))
    _external(perm)
    for i in range(16):
        ev.add_constraint(is_first * (perm[i] - out_state[i]))
    full = list(in_state)
    for i in range(16):
        full[i] = full[i] + rc0[i]
    full = [_pow5(x) for x in full]
    for i in range(16):
        ev.add_constraint(is_full * (mid[i] - full[i]))
        full[i] = mid[i]
    _external(full)
    for i in range(16):
        full[i] = full[i] + rc1[i]
    full = [_pow5(x) for x in full]
    _external(full)
    for i in range(16):
        ev.add_constraint(is_full * (out_state[i] - full[i]))
    part = list(in_state)
    for r in range(14):
        part[0] = part[0] + rc0[r]
        part[0] = _pow5(part[0])
        ev.add_constraint(is_partial * (mid[r] - part[0]))
        part[0] = mid[r]
        _internal(part)
    for i in range(16):
        ev.add_constraint(is_partial * (out_state[i] - part[i]))
    in_left_id = round_id + round_id
    in_right_id = in_left_id + o
    out_left_id = in_right_id + o
    out_right_id = out_left_id + o
    sel = ext1_nz * is_first
    idv = (is_first * ext1) + (not_first * in_left_id)
    a, b = combine_ef(in_state[0:4]), combine_ef(in_state[4:8])
    ev.add_to_relation(rel, sel - not_first, [idv, a, b])
    sel = ext2_nz * is_first
    idv = (is_first * ext2) + (not_first * in_right_id)
    a, b = combine_ef(in_state[8:12]), combine_ef(in_state[12:16])
    ev.add_to_relation(rel, sel - not_first, [idv, a, b])
    sel = ext1_nz * is_last
    idv = (is_last * ext1) + (not_last * out_left_id)
    a, b = combine_ef(out_state[0:4]), combine_ef(out_state[4:8])
    ev.add_to_relation(rel, sel + not_last, [idv, a, b])
    sel = ext2_nz * is_last
    idv = (is_last * ext2) + (not_last * out_right_id)
    a, b = combine_ef(out_state[8:12]), combine_ef(out_state[12:16])
    ev.add_to_relation(rel, sel + not_last, [idv, a, b])
    ev.add_to_relation(rel, is_first * not_last, [swap_bit, swap_bit_addr])
    ev.finalize_logup(3)


def composition_check(cs, pv, shape, fs):
    """CompositionCheck::compute (components/recursive/composition/src/lib.rs:33-121)"""
    sv = pv.sampled_values
    acc = [qm31_zero(cs), fs["random_coeff"]]
    sp = shape.split
    for comp, lo, evaluate in (("plonk", 0, evaluate_plonk), ("poseidon", 1, evaluate_poseidon)):
        mask = [sv[t][:sp[t]] if lo == 0 else sv[t][sp[t]:] for t in range(3)]
        log_size = pv.p.log_size_plonk if lo == 0 else pv.p.log_size_poseidon
        total = pv.plonk_total_sum if lo == 0 else pv.poseidon_total_sum
        # argument order of EvalAtRowVar::new(mask, total_sum, coset_vanishing(..).inv(), log, acc): the inverse is
        # evaluated before LogupAtRowVar::new computes cumsum_shift
        dinv = qm31_inv(coset_vanishing(fs["oods_point"], log_size))
        ev = EvalAtRow(mask, total, dinv, log_size, acc)
        evaluate(cs, fs, ev)
    computed = acc[0]
    left = ((sv[3][0][0] + sv[3][1][0].shift_by_i()) + sv[3][2][0].shift_by_j()) + sv[3][3][0].shift_by_ij()
    right = ((sv[3][4][0] + sv[3][5][0].shift_by_i()) + sv[3][6][0].shift_by_j()) + sv[3][7][0].shift_by_ij()
    expected = left + (right * fs["oods_point"].repeated_double_x_only(shape.composition_log_degree_bound - 2))
    computed.equalverify(expected)
    return computed


def single_path_verify(cs, proof, sibling_hashes, columns, root, bits):
    """SinglePathMerkleProofVar::verify (components/recursive/data_structures/src/lib.rs:315-354)"""
    assert bits.get_value() == proof.query
    cur = hash_m31_columns_get_rate(columns[proof.depth])
    for i in range(proof.depth):
        h = proof.depth - i - 1
        if h in columns:
            column_hash = hash_m31_columns_get_capacity(columns[h])
            cur = permute_get_rate(cur, sibling_hashes[i], (bits.value[i], bits.variables[i]))
            cur = permute_get_rate(cur, column_hash)
        else:
            cur = permute_get_rate(cur, sibling_hashes[i], (bits.value[i], bits.variables[i]))
    assert cur.value == root.value, "merkle root"
    cur.equalverify(root)


class PairVar:
    """SinglePairMerkleProofVar (components/recursive/data_structures/src/lib.rs:358-464)"""
    def __init__(self, cs, proof):
        self.cs, self.value = cs, proof
        self.sibling_hashes = [Half.single_use_witness_only(cs, s) for s in proof.sibling_hashes]
        self.self_columns = {k: qm31_witness(cs, v) for k, v in sorted(proof.self_columns.items())}
        self.siblings_columns = {k: qm31_witness(cs, v) for k, v in sorted(proof.siblings_columns.items())}

    def verify(self, root, bits):
        cs, depth = self.cs, self.value.depth
        assert bits.get_value() == self.value.query
        self_hash = hash_qm31_columns_get_rate([self.self_columns[depth], qm31_zero(cs)])
        sibling_hash = hash_qm31_columns_get_rate([self.siblings_columns[depth], qm31_zero(cs)])
        for i in range(depth):
            h = depth - i - 1
            sw = (bits.value[i], bits.variables[i])
            if h not in self.self_columns:
                self_hash = permute_get_rate(self_hash, sibling_hash, sw)
                if i != depth - 1:
                    sibling_hash = self.sibling_hashes[i]
            else:
                self_column_hash = hash_qm31_columns_get_capacity([self.self_columns[h], qm31_zero(cs)])
                sibling_column_hash = hash_qm31_columns_get_capacity([self.siblings_columns[h], qm31_zero(cs)])
                self_hash = permute_get_rate(self_hash, sibling_hash, sw)
                self_hash = permute_get_rate(self_hash, self_column_hash)
                sibling_hash = permute_get_rate(self.sibling_hashes[i], sibling_column_hash)
        assert self_hash.value == root.value, "pair merkle root"
        self_hash.equalverify(root)


def answers(cs, pv, shape, fs, hints, oods_witness, last_decommit=None):
    """AnswerResults::compute (components/recursive/answer/src/lib.rs:34-354)"""
    nq = shape.n_queries
    # shifted mask points: shift sets in first-appearance order (0, -1); Plonk before Poseidon (see module docstring)
    shifted = {}
    for comp, log in (("plonk", shape.log_plonk), ("poseidon", shape.log_poseidon)):
        step = cp_gen(log)
        for s in (0, -1):
            shifted[(comp, s)] = oods_witness.add_m31_point((1, 0) if s == 0 else cp_neg(step))
    # samples per column, flatten order: (shift key, point, value)
    samples = []                                                         # [(log_size, [(key, point, value)])]
    for t in range(4):
        for c, col in enumerate(pv.sampled_values[t]):
            comp = shape.column_component[t][c]
            entries = []
            if t == 0 or comp is None:                                   # mask_points[PREPROCESSED] and composition: (Zero, oods)
                assert len(col) == 1
                entries.append(("zero", oods_witness, col[0]))
            else:
                shifts = [0] if len(col) == 1 else [-1, 0]
                comp_log = shape.log_plonk if comp == "plonk" else shape.log_poseidon
                for s, v in zip(shifts, col):
                    key = "zero" if s == 0 else (s, comp_log)            # ShiftIndex::from_shift
                    entries.append((key, shifted[(comp, s)], v))
            samples.append((shape.column_log_sizes[t][c], entries))
    lo = (shape.blowup + 1) if last_decommit is not None else (shape.log_last + shape.blowup + 1)
    qpos = query_positions_per_log_size(lo, shape.max_first, fs["raw_queries"])
    if len({q.bits.get_value() for q in qpos[shape.max_first]}) != nq:
        raise NotImplementedError("duplicated queries at the largest size (answer/src/lib.rs:190-195)")
    tree_depth = [max(shape.log_plonk, shape.log_poseidon) + shape.blowup] * 3 + [shape.max_first]
    if last_decommit is not None:
        dec = last_decommit(tree_depth, qpos)                            # LastDecommitVar::compute
    else:
        # DecommitmentVar::new: per tree, per query: sibling hashes (values only), column witnesses in ascending layer order
        dec = []
        for t in range(4):
            per_q = []
            for i in range(nq):
                sp = SinglePath(hints, t, i, qpos[tree_depth[t]][i].bits.get_value())
                sib = [Half.single_use_witness_only(cs, s) for s in sp.sibling_hashes]
                cols = {k: [m31_witness(cs, v) for v in vs] for k, vs in sorted(sp.columns.items())}
                per_q.append((sp, sib, cols))
            dec.append(per_q)
        for t in range(4):
            for i in range(nq):
                sp, sib, cols = dec[t][i]
                single_path_verify(cs, sp, sib, cols, pv.commitments[t], qpos[tree_depth[t]][i].bits)
    queried = {}
    for L in shape.all_log_sizes:
        queried[L] = [[v for t in range(4) for v in dec[t][i][2].get(L, [])] for i in range(nq)]
    fri_answers, domain_points = {}, {}
    for L in sorted(shape.all_log_sizes, reverse=True):
        cols = [e for (log, e) in samples if log == L]
        # ColumnSampleBatchVar::new_vec (answer/src/data_structures.rs:42-64)
        order, groups = [], {}
        for ci, entries in enumerate(cols):
            for key, point, value in entries:
                if key not in groups:
                    groups[key] = []
                    order.append(key)
                groups[key].append((point, ci, value))
        batches = [(groups[k][0][0], [(ci, v) for (_, ci, v) in groups[k]]) for k in order]
        # column_line_coeffs_var (:162-189) + complex_conjugate_line_coeffs_var (:137-160)
        alpha = qm31_constant(cs, (0, 0, P - 2, 0))
        line_coeffs = []
        for point, cvs in batches:
            lc = []
            for _, sv in cvs:
                value0, value1 = qm31_decompose_cm31(sv)
                y0, y1 = qm31_decompose_cm31(point.y)
                a, c = value1, y1
                b = (value0 * y1) - (value1 * y0)
                lc.append((alpha * a, alpha * b, alpha * c))
                alpha = alpha * fs["after"]
            line_coeffs.append(lc)
        ans, dps = [], []
        for i in range(nq):
            q = qpos[L][i]
            dp = q.get_next_point()
            # denominator_inverses_var (:103-126)
            dinv = []
            for point, _ in batches:
                prx, pix = qm31_decompose_cm31(point.x)
                pry, piy = qm31_decompose_cm31(point.y)
                a = prx - dp.x
                a = a * piy
                b = pry - dp.y
                b = b * pix
                dinv.append(cm31_inv(a - b))
            # accumulate_row_quotients_var (:70-101)
            row_acc = qm31_zero(cs)
            for (point, cvs), lc, di in zip(batches, line_coeffs, dinv):
                num = qm31_zero(cs)
                for (ci, _), (a, b, c) in zip(cvs, lc):
                    value = queried[L][i][ci] * c
                    linear = (a * dp.y) + b
                    num = num + (value - linear)
                row_acc = row_acc + (num * di)
            ans.append(row_acc)
            dps.append(dp)
        fri_answers[L], domain_points[L] = ans, dps
    return dict(qpos=qpos, fri_answers=fri_answers, domain_points=domain_points)


def folding(cs, pv, shape, fs, ans, hints):
    """FoldingResults::compute (components/recursive/folding/src/lib.rs:12-205)"""
    nq, qpos = shape.n_queries, ans["qpos"]
    proofs = []
    for i in range(nq):
        bits = qpos[shape.max_first][i].bits
        pr = PairVar(cs, SinglePair(hints, 0, i, bits.get_value()))
        pr.verify(pv.first_layer_commitment, bits)
        proofs.append(pr)
    for L in sorted(shape.all_log_sizes, reverse=True):
        for i in range(nq):
            proofs[i].self_columns[L].equalverify(ans["fri_answers"][L][i])
    folded_results = {}
    for L in shape.all_log_sizes:
        out = []
        for pr, q in zip(proofs, qpos[L]):
            self_val, sibling_val = pr.self_columns[L], pr.siblings_columns[L]
            point = q.point.double()
            y_inv = m31_inv(point.y)
            l, r = qm31_swap(self_val, sibling_val, q.bits.value[0], q.bits.variables[0])
            nl = l + r
            nr = (l - r) * y_inv
            out.append(nl + (nr * fs["fri_alphas"][shape.max_first - L]))
        folded_results[L] = out
    log_size = shape.max_first
    folded = [qm31_zero(cs) for _ in range(nq)]
    for i in range(len(pv.inner_layer_commitments)):
        if log_size in folded_results:
            a = fs["fri_alphas"][i]
            a = a * a
            folded = [(a * v) + b for v, b in zip(folded, folded_results[log_size])]
        log_size -= 1
        new_folded = []
        for k in range(nq):
            q = qpos[log_size][k]
            mp = PairVar(cs, SinglePair(hints, 1 + i, k, q.bits.get_value()))
            self_val, sibling_val = mp.self_columns[log_size], mp.siblings_columns[log_size]
            folded[k].equalverify(self_val)
            x_inv = m31_inv(q.point.x)
            l, r = qm31_swap(self_val, sibling_val, q.bits.value[0], q.bits.variables[0])
            nl = l + r
            nr = (l - r) * x_inv
            new_folded.append(nl + (nr * fs["fri_alphas"][i + 1]))
            mp.verify(pv.inner_layer_commitments[i], q.bits)
        folded = new_folded
    for q, v in zip(qpos[log_size], folded):
        if len(pv.last_poly) == 1:
            v.equalverify(pv.last_poly[0])
        else:
            x = q.get_next_point_x()
            v.equalverify(line_poly_eval_at_point(cs, pv.last_poly, x))
    return folded


INPUTS_SINGLE = [(1, Q1)]                                                # examples/single-proof/src/main.rs:33
INPUTS_RECURSIVE = [(1, Q1), (2, (0, 1, 0, 0)), (3, (0, 0, 1, 0))]       # examples/multi-proofs/src/main.rs:52-59


def verifier_circuit(blob, inputs, multipliers=1, verify_out_cls=None, finalize=True):
    """examples/single-proof/src/main.rs:33-90 / examples/multi-proofs/src/main.rs:49-139: returns the CS after
    pad / check_arithmetics / populate_logup_arguments / check_poseidon_invocations, plus the C oracle's VerifyOut."""
    p = parse_proof(blob)
    shape = Shape(p)
    out, hints = compute_hints(blob, ([i for i, _ in inputs], [list(v) for _, v in inputs]), verify_out_cls)
    assert out.verdict == 0, "the reference panics on a rejected proof (stage %d)" % out.stage
    cs = CS()
    marks = []
    for _ in range(multipliers):
        pv = ProofVar(cs, p)
        marks.append(("alloc", len(cs.a_wire), len(cs.flow)))
        const_in = [(i, qm31_constant(cs, v)) for i, v in inputs]
        fs = fiat_shamir(cs, pv, shape, const_in)
        marks.append(("fiat_shamir", len(cs.a_wire), len(cs.flow)))
        composition_check(cs, pv, shape, fs)
        marks.append(("composition", len(cs.a_wire), len(cs.flow)))
        oods_w = PointQM31(qm31_witness(cs, tuple(out.oods_x)), qm31_witness(cs, tuple(out.oods_y)))
        ans = answers(cs, pv, shape, fs, hints, oods_w)
        marks.append(("answer", len(cs.a_wire), len(cs.flow)))
        folding(cs, pv, shape, fs, ans, hints)
        marks.append(("folding", len(cs.a_wire), len(cs.flow)))
        cs.last = dict(fs=fs, ans=ans, shape=shape)
    cs.marks = marks
    if finalize:
        cs.pad()
        bad = cs.check_arithmetics()
        assert bad < 0, "check_arithmetics fails at row %d" % bad
        cs.populate_logup_arguments()
        bad = cs.check_poseidon_invocations()
        assert bad < 0, "check_poseidon_invocations fails at entry %d" % bad
    return cs, out


class FriOnlyShape:
    """the part of Shape the folding stage reads, from the seven shape words of a synthetic instance"""
    def __init__(self, words):
        self.log_plonk, self.log_poseidon, self.pow_bits, self.blowup, self.log_last, self.n_queries, n_inner = [int(x) for x in words[1:8]]
        self.n_inner = n_inner
        self.max_first = self.log_last + self.blowup + 1 + n_inner
        self.all_log_sizes = sorted({self.log_plonk + self.blowup, self.log_poseidon + self.blowup, self.max_first})


def folding_circuit(words, verify_out_cls, finalize=True):
    """The folding stage alone (components/recursive/folding/src/lib.rs:12-205) over a synthetic FRI + Merkle instance (BASELINE
    configs[4] part i: "fri_answers supplied as witnesses"): commitments, last-layer polynomial, alphas, query positions and first-layer
    answers are witnesses, in that order; then `folding` as in the verifier circuit.  words: the instance blob (uint32)."""
    lib = _orc()
    w = np.ascontiguousarray(words, dtype=np.uint32)
    out, hints = verify_out_cls(), Hints()
    lib.orc_fri_verify_synth_hints(w.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(w.size), ctypes.byref(out), ctypes.byref(hints))
    assert out.verdict == 0, "the reference panics on a rejected instance (stage %d)" % out.stage
    shape = FriOnlyShape(w)
    nq, n_inner = shape.n_queries, shape.n_inner
    cs = CS()

    class _P:
        pass
    pv = _P()
    pv.first_layer_commitment = Half.new_witness(cs, [int(x) for x in w[w[80]: w[80] + 8]])
    pv.inner_layer_commitments = [Half.new_witness(cs, [int(x) for x in w[w[96 + i]: w[96 + i] + 8]]) for i in range(n_inner)]
    pv.last_poly = [qm31_witness(cs, tuple(int(x) for x in w[w[81] + 4 * k: w[81] + 4 * k + 4])) for k in range(1 << shape.log_last)]
    fs = {"fri_alphas": [qm31_witness(cs, tuple(out.fri_alphas[l])) for l in range(n_inner + 1)]}
    mask = (1 << shape.max_first) - 1
    fs["raw_queries"] = [m31_witness(cs, int(out.raw_queries[i]) & mask) for i in range(nq)]
    qpos = query_positions_per_log_size(shape.log_last + shape.blowup + 1, shape.max_first, fs["raw_queries"])
    fri_answers = {}
    for g, L in enumerate(sorted(shape.all_log_sizes, reverse=True)):
        assert int(out.log_sizes[g]) == L
        fri_answers[L] = [qm31_witness(cs, tuple(out.fri_answers[g][i])) for i in range(nq)]
    folding(cs, pv, shape, fs, {"qpos": qpos, "fri_answers": fri_answers}, hints)
    if finalize:
        cs.pad()
        bad = cs.check_arithmetics()
        assert bad < 0, "check_arithmetics fails at row %d" % bad
        cs.populate_logup_arguments()
        bad = cs.check_poseidon_invocations()
        assert bad < 0, "check_poseidon_invocations fails at entry %d" % bad
    return cs, out


# ---- the last-layer circuit (components/last/*, examples/last-layer/src/main.rs:26-94) --------------------------------
def native_hash_rate(words):
    """Poseidon31MerkleHasher::hash_column_get_capacity then permute_get_rate([0; 8] || capacity)
    (last/fiat_shamir/src/lib.rs:46-54, last/answer/src/data_structures/merkle_proofs.rs:190-195)"""
    arr = np.ascontiguousarray(words, dtype=np.uint32)
    cap = np.zeros(8, dtype=np.uint32)
    _orc().orc_hash_column_get_capacity(arr.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(arr.size), cap.ctypes.data_as(ctypes.c_void_p))
    return poseidon2_permute([0] * 8 + [int(x) for x in cap])[0:8]


def pack_columns(values):
    """LastSinglePathMerkleProofInput::from_proof (merkle_proofs.rs:170-206): <= 4 values one QM31, <= 8 two, else their hash"""
    if len(values) <= 4:
        v = list(values) + [0] * (4 - len(values))
        return [tuple(v)]
    if len(values) <= 8:
        v = list(values) + [0] * (8 - len(values))
        return [tuple(v[0:4]), tuple(v[4:8])]
    h = native_hash_rate(values)
    return [tuple(h[0:4]), tuple(h[4:8])]


def last_layer_circuit(blob, verify_out_cls, finalize=True):
    """examples/last-layer/src/main.rs:26-97 on a Poseidon31 proof (the hybrid SHA-256 channel of hybrid_hash.bin is
    parity-unpinned; the circuit itself never replays the channel -- every Fiat-Shamir value is a public input)."""
    p = parse_proof(blob)
    shape = Shape(p)
    out, hints = compute_hints(blob, ([i for i, _ in INPUTS_RECURSIVE], [list(v) for _, v in INPUTS_RECURSIVE]), verify_out_cls)
    assert out.verdict == 0
    nq = shape.n_queries
    cs = CSWithout()
    tree_depth = [max(shape.log_plonk, shape.log_poseidon) + shape.blowup] * 3 + [shape.max_first]
    pos_max = [int(out.query_pos[0][i]) for i in range(nq)]
    # ---- public inputs, in the order of main.rs:62-69 --------------------------------------------------------------
    flat_words = [w for tree in p.sampled_values for col in tree for v in col for w in v]
    inp = {}
    inp["t"] = qm31_input(cs, tuple(out.oods_t))                        # LastFiatShamirInputVar (last/fiat_shamir/src/lib.rs:104-150)
    inp["sampled_values_hash"] = Half.new_variables(cs, native_hash_rate(flat_words), "input")
    inp["plonk_total_sum"] = qm31_input(cs, p.plonk_total_sum)
    inp["poseidon_total_sum"] = qm31_input(cs, p.poseidon_total_sum)
    inp["z"] = qm31_input(cs, tuple(out.z))
    inp["alpha"] = qm31_input(cs, tuple(out.alpha))
    inp["random_coeff"] = qm31_input(cs, tuple(out.random_coeff))
    inp["after"] = qm31_input(cs, tuple(out.after_coeff))
    inp["packed_queries"] = [qm31_input(cs, tuple((pos_max[k:k + 4] + [0, 0, 0])[:4])) for k in range(0, nq, 4)]
    inp["fri_alphas"] = [qm31_input(cs, tuple(out.fri_alphas[k])) for k in range(len(p.inner_layers) + 1)]
    paths = [[SinglePath(hints, t, i, None) for i in range(nq)] for t in range(4)]
    dec_in = [[{k: [qm31_input(cs, q) for q in pack_columns(vs)] for k, vs in sorted(paths[t][i].columns.items())} for i in range(nq)]
              for t in range(4)]                                         # LastDecommitInputVar
    pairs0 = [SinglePair(hints, 0, i, None) for i in range(nq)]

    def pair_input(sp):
        self_c = {k: qm31_input(cs, v) for k, v in sorted(sp.self_columns.items())}
        sib_c = {k: qm31_input(cs, v) for k, v in sorted(sp.siblings_columns.items())}
        return self_c, sib_c
    first_in = [pair_input(sp) for sp in pairs0]                         # LastFirstLayerInputVar
    inner_in = {}                                                        # LastInnerLayersInputVar: BTreeMap, ascending log size
    n_inner = len(p.inner_layers)
    for li in range(n_inner - 1, -1, -1):
        inner_in[shape.max_first - 1 - li] = [pair_input(SinglePair(hints, 1 + li, k, None)) for k in range(nq)]
    n_public = cs.num_input
    # ---- LastPlonkWithPoseidonProofVar::new_witness (last/data_structures/src/lib.rs:28-82) ---------------------------
    class PV:
        pass
    pv = PV()
    pv.p = p
    pv.log_size_plonk = m31_witness(cs, p.log_size_plonk)
    pv.log_size_poseidon = m31_witness(cs, p.log_size_poseidon)
    pv.plonk_total_sum = qm31_witness(cs, p.plonk_total_sum)
    pv.poseidon_total_sum = qm31_witness(cs, p.poseidon_total_sum)
    pv.sampled_values = [[[qm31_witness(cs, v) for v in col] for col in tree] for tree in p.sampled_values]
    pv.last_poly = [qm31_witness(cs, c) for c in p.last_coeffs]
    marks = [("alloc", len(cs.a_wire))]
    # ---- LastFiatShamirResults::compute (last/fiat_shamir/src/lib.rs:164-216) ---------------------------------------------
    oods_point = PointQM31.from_t(inp["t"])
    flat = [v for tree in pv.sampled_values for col in tree for v in col]
    hash_qm31_columns_get_rate(flat).equalverify(inp["sampled_values_hash"])
    z, alpha = inp["z"], inp["alpha"]
    alpha_powers = [qm31_one(cs), alpha, alpha * alpha]
    queries = []
    for packed in inp["packed_queries"]:
        queries += qm31_decompose_m31(packed)
    queries = queries[:nq]
    input_sum = qm31_zero(cs)
    s1 = (qm31_one(cs) + alpha) - z
    input_sum = input_sum + qm31_inv(s1)
    alpha_two = alpha + alpha
    s2 = (V(cs, (0, 1, 0, 0), 2, 2) + alpha_two) - z
    input_sum = input_sum + qm31_inv(s2)
    alpha_three = alpha_two + alpha
    s3 = (V(cs, (0, 0, 1, 0), 3, 2) + alpha_three) - z
    input_sum = input_sum + qm31_inv(s3)
    ((input_sum + inp["poseidon_total_sum"]) + inp["plonk_total_sum"]).equalverify(qm31_zero(cs))
    fs = dict(z=z, alpha=alpha, alpha_powers=alpha_powers, random_coeff=inp["random_coeff"], after=inp["after"], oods_point=oods_point,
              raw_queries=queries, fri_alphas=inp["fri_alphas"])
    marks.append(("fiat_shamir", len(cs.a_wire)))

    # ---- LastAnswerResults::compute (last/answer/src/lib.rs:30-262) ----------------------------------------------------------
    def last_decommit(depths, qpos):
        """LastDecommitVar::compute / LastSinglePathMerkleProofVar::from_proof_and_input (merkle_proofs.rs:114-160)"""
        dec = []
        for t in range(4):
            per_q = []
            for i in range(nq):
                cols = {}
                for L, vs in sorted(paths[t][i].columns.items()):
                    vars_ = [m31_witness(cs, v) for v in vs]
                    packed = dec_in[t][i][L]
                    if len(vars_) <= 8:
                        for k in range(0, len(vars_), 4):
                            d = qm31_decompose_m31(packed[k // 4])
                            for l, r in zip(vars_[k:k + 4], d):
                                l.equalverify(r)
                    else:
                        h = hash_m31_columns_get_rate(vars_).to_qm31()
                        h[0].equalverify(packed[0])
                        h[1].equalverify(packed[1])
                    cols[L] = vars_
                per_q.append((None, None, cols))
            dec.append(per_q)
        return dec
    ans = answers(cs, pv, shape, fs, hints, oods_point, last_decommit)
    marks.append(("answer", len(cs.a_wire)))
    # ---- LastFoldingResults::compute (last/folding/src/lib.rs:14-161) ----------------------------------------------------
    qpos = ans["qpos"]
    for L in sorted(shape.all_log_sizes, reverse=True):
        for i in range(nq):
            first_in[i][0][L].equalverify(ans["fri_answers"][L][i])
    folded_results = {}
    for L in shape.all_log_sizes:
        res = []
        for (self_c, sib_c), q in zip(first_in, qpos[L]):
            point = q.point.double()
            y_inv = m31_inv(point.y)
            l, r = qm31_swap(self_c[L], sib_c[L], q.bits.value[0], q.bits.variables[0])
            nl = l + r
            nr = (l - r) * y_inv
            res.append(nl + (nr * fs["fri_alphas"][shape.max_first - L]))
        folded_results[L] = res
    log_size = shape.max_first
    folded = [qm31_zero(cs) for _ in range(nq)]
    for i in range(n_inner):
        if log_size in folded_results:
            a = fs["fri_alphas"][i]
            a = a * a
            folded = [(a * v) + b for v, b in zip(folded, folded_results[log_size])]
        log_size -= 1
        new_folded = []
        for k in range(nq):
            q = qpos[log_size][k]
            self_c, sib_c = inner_in[log_size][k]
            folded[k].equalverify(self_c[log_size])
            x_inv = m31_inv(q.point.x)
            l, r = qm31_swap(self_c[log_size], sib_c[log_size], q.bits.value[0], q.bits.variables[0])
            nl = l + r
            nr = (l - r) * x_inv
            new_folded.append(nl + (nr * fs["fri_alphas"][i + 1]))
        folded = new_folded
    for q, v in zip(qpos[log_size], folded):
        if len(pv.last_poly) == 1:
            v.equalverify(pv.last_poly[0])
        else:
            v.equalverify(line_poly_eval_at_point(cs, pv.last_poly, q.get_next_point_x()))
    marks.append(("folding", len(cs.a_wire)))
    cs.marks, cs.n_public_inputs = marks, n_public
    if finalize:
        cs.pad()
        bad = cs.check_arithmetics()
        assert bad < 0, "check_arithmetics fails at row %d" % bad
        cs.populate_logup_arguments()
    return cs, out


def value_log_arrays(cs):
    """(ops [n, 8], perms [m, 32]) uint32: the value log in the layout oracle/orc_tape.c replays"""
    return np.array(cs.log, dtype=np.uint32).reshape(-1, 8), np.array(cs.perm_log, dtype=np.uint32).reshape(-1, 32)


def trace_digest(cols):
    return hashlib.sha256(np.ascontiguousarray(cols, dtype="<u4").tobytes()).hexdigest()
