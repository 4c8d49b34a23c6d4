"""TEST INFRASTRUCTURE: runs the reference's generated EVM verifier program (release-v0.13.1/evm_verifier.yul, the Yul source
of evm_verifier.bin) on a calldata blob, statement by statement -- the `verifier bytecode check` of BASELINE configs[4]
(/root/reference/integration/tests/e2e_tests.rs:185-199 -> EVMVerifier::verify_evm_proof), without an EVM.

The program is straight-line Yul over a small builtin set (mload/mstore/mstore8/calldataload, addmod/mulmod/mod/add/sub/shl,
and/or/not/eq/lt, keccak256, staticcall to the precompiles 0x5 modexp, 0x6 ecAdd, 0x7 ecMul, 0x8 ecPairing, revert/return) plus
one helper function (validate_ec_point).  This interpreter parses exactly that subset; the elliptic-curve precompiles are
pluggable so that the same run can be made with the independent big-integer model (tests/pairing_model.py) and with the
PRODUCT's host-side curve / pairing code (scroll-prover_b200/pairing_bn254.hpp + csrc/ec.cuh through tests/host_emul).
"""
from __future__ import annotations

import re

M256 = (1 << 256) - 1


# ------------------------------------------------------------------------------------------------ Keccak-256 (pre-SHA3 padding)
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
       0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
       0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
       0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def _rol(x, n):
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def _keccak_f(a):
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    return a


def keccak256(data: bytes) -> bytes:
    rate = 136
    p = bytearray(data)
    p.append(0x01)
    while len(p) % rate:
        p.append(0)
    p[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(p), rate):
        blk = p[off:off + rate]
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(blk[8 * i:8 * i + 8], "little")
        a = _keccak_f(a)
    out = b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))
    return out


# ------------------------------------------------------------------------------------------------ a parser for the Yul subset
_TOK = re.compile(r"\s*(0x[0-9a-fA-F]+|\d+|[A-Za-z_][A-Za-z_0-9]*|:=|->|[(){},:]|\"[^\"]*\")")


def _tokens(src: str):
    pos, out = 0, []
    src = re.sub(r"//[^\n]*", "", src)
    while pos < len(src):
        m = _TOK.match(src, pos)
        if not m:
            if src[pos:].strip() == "":
                break
            raise SyntaxError(f"yul: cannot tokenise at {src[pos:pos + 40]!r}")
        out.append(m.group(1))
        pos = m.end()
    return out


class _Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def take(self, want=None):
        tok = self.t[self.i]
        if want is not None and tok != want:
            raise SyntaxError(f"yul: expected {want!r}, got {tok!r} at token {self.i}")
        self.i += 1
        return tok

    def expr(self):
        tok = self.take()
        if tok.startswith("0x"):
            return ("lit", int(tok, 16))
        if tok.isdigit():
            return ("lit", int(tok))
        if tok == "true":
            return ("lit", 1)
        if tok == "false":
            return ("lit", 0)
        if self.peek() == "(":
            self.take("(")
            args = []
            while self.peek() != ")":
                args.append(self.expr())
                if self.peek() == ",":
                    self.take(",")
            self.take(")")
            return ("call", tok, args)
        return ("var", tok)

    def typed_name(self):
        name = self.take()
        if self.peek() == ":":  # `name:bool`
            self.take(":")
            self.take()
        return name

    def block(self):
        self.take("{")
        stmts = []
        while self.peek() != "}":
            stmts.append(self.stmt())
        self.take("}")
        return ("block", stmts)

    def stmt(self):
        tok = self.peek()
        if tok == "{":
            return self.block()
        if tok == "let":
            self.take()
            name = self.typed_name()
            if self.peek() != ":=":
                return ("let", name, ("lit", 0))  # `let v` without an initial value is zero
            self.take(":=")
            return ("let", name, self.expr())
        if tok == "if":
            self.take()
            cond = self.expr()
            return ("if", cond, self.block())
        if tok == "function":
            self.take()
            name = self.take()
            self.take("(")
            params = []
            while self.peek() != ")":
                params.append(self.typed_name())
                if self.peek() == ",":
                    self.take(",")
            self.take(")")
            rets = []
            if self.peek() == "->":
                self.take()
                rets.append(self.typed_name())
            return ("function", name, params, rets, self.block())
        if self.t[self.i + 1] == ":=":
            name = self.take()
            self.take(":=")
            return ("assign", name, self.expr())
        return ("expr", self.expr())


def parse_runtime(yul_source: str):
    """the statements of `object "Runtime" { code { ... } }`"""
    start = yul_source.index('object "Runtime"')
    body = yul_source[yul_source.index("code", start) + 4:]
    p = _Parser(_tokens(body))
    return p.block()


# ------------------------------------------------------------------------------------------------ the machine
class Halt(Exception):
    def __init__(self, reverted):
        self.reverted = reverted


class Machine:
    def __init__(self, calldata: bytes, ec_add, ec_mul, ec_pairing):
        self.calldata, self.mem = calldata, bytearray()
        self.ec_add, self.ec_mul, self.ec_pairing = ec_add, ec_mul, ec_pairing
        self.funcs, self.precompile_calls = {}, {5: 0, 6: 0, 7: 0, 8: 0}
        self.keccak_calls = 0

    # memory
    def _grow(self, end):
        if end > len(self.mem):
            self.mem.extend(bytes(end - len(self.mem)))

    def mload(self, p):
        self._grow(p + 32)
        return int.from_bytes(self.mem[p:p + 32], "big")

    def mstore(self, p, v):
        self._grow(p + 32)
        self.mem[p:p + 32] = (v & M256).to_bytes(32, "big")

    def mread(self, p, n):
        self._grow(p + n)
        return bytes(self.mem[p:p + n])

    def staticcall(self, addr, in_p, in_n, out_p, out_n):
        data = self.mread(in_p, in_n)
        self.precompile_calls[addr] = self.precompile_calls.get(addr, 0) + 1
        w = lambda i: int.from_bytes(data[32 * i:32 * i + 32], "big")
        try:
            if addr == 5:  # modexp: <len_b, len_e, len_m, b, e, m>
                lb, le, lm = w(0), w(1), w(2)
                b = int.from_bytes(data[96:96 + lb], "big")
                e = int.from_bytes(data[96 + lb:96 + lb + le], "big")
                m = int.from_bytes(data[96 + lb + le:96 + lb + le + lm], "big")
                out = (pow(b, e, m) if m else 0).to_bytes(lm, "big")
            elif addr == 6:
                out = self.ec_add((w(0), w(1)), (w(2), w(3)))
            elif addr == 7:
                out = self.ec_mul((w(0), w(1)), w(2))
            elif addr == 8:
                if in_n % 192:
                    return 0
                out = self.ec_pairing(data)
            else:
                return 0
        except ValueError:
            return 0  # the precompile rejects its input
        if out is None:
            return 0
        self._grow(out_p + out_n)
        self.mem[out_p:out_p + out_n] = out[:out_n]
        return 1

    # evaluation
    def ev(self, e, env):
        kind = e[0]
        if kind == "lit":
            return e[1]
        if kind == "var":
            for scope in reversed(env):
                if e[1] in scope:
                    return scope[e[1]]
            raise NameError(e[1])
        name, args = e[1], e[2]
        if name in self.funcs:
            params, rets, body = self.funcs[name]
            scope = {p: self.ev(a, env) for p, a in zip(params, args)}
            for r in rets:
                scope[r] = 0
            self.run(body, [env[0], scope])
            return scope[rets[0]] if rets else 0
        a = [self.ev(x, env) for x in args]
        if name == "mload": return self.mload(a[0])
        if name == "mstore": self.mstore(a[0], a[1]); return 0
        if name == "mstore8": self._grow(a[0] + 1); self.mem[a[0]] = a[1] & 0xFF; return 0
        if name == "calldataload": return int.from_bytes((self.calldata[a[0]:a[0] + 32] + bytes(32))[:32], "big")
        if name == "mulmod": return (a[0] * a[1]) % a[2] if a[2] else 0
        if name == "addmod": return (a[0] + a[1]) % a[2] if a[2] else 0
        if name == "mod": return a[0] % a[1] if a[1] else 0
        if name == "add": return (a[0] + a[1]) & M256
        if name == "sub": return (a[0] - a[1]) & M256
        if name == "shl": return (a[1] << a[0]) & M256 if a[0] < 256 else 0
        if name == "and": return a[0] & a[1]
        if name == "or": return a[0] | a[1]
        if name == "not": return (~a[0]) & M256 if a[0] > 1 else 1 - a[0]  # the program only negates booleans
        if name == "eq": return int(a[0] == a[1])
        if name == "lt": return int(a[0] < a[1])
        if name == "gas": return M256
        if name == "keccak256": self.keccak_calls += 1; return int.from_bytes(keccak256(self.mread(a[0], a[1])), "big")
        if name == "staticcall": return self.staticcall(a[1], a[2], a[3], a[4], a[5])
        if name == "revert": raise Halt(True)
        if name == "return": raise Halt(False)
        raise NotImplementedError(f"yul builtin {name}")

    def run(self, block, env):
        scope = {}
        env = env + [scope]
        for s in block[1]:
            k = s[0]
            if k == "function":
                self.funcs[s[1]] = (s[2], s[3], s[4])
        for s in block[1]:
            k = s[0]
            if k == "let":
                scope[s[1]] = self.ev(s[2], env)
            elif k == "assign":
                for sc in reversed(env):
                    if s[1] in sc:
                        sc[s[1]] = self.ev(s[2], env)
                        break
                else:
                    raise NameError(s[1])
            elif k == "expr":
                self.ev(s[1], env)
            elif k == "if":
                if self.ev(s[1], env):
                    self.run(s[2], env)
            elif k == "block":
                self.run(s, env)


def verify(yul_source: str, calldata: bytes, ec_add, ec_mul, ec_pairing):
    """True iff the program returns without reverting; also returns the machine (precompile / hash counts)."""
    m = Machine(calldata, ec_add, ec_mul, ec_pairing)
    try:
        m.run(parse_runtime(yul_source), [{}])
    except Halt as h:
        return (not h.reverted), m
    return True, m
