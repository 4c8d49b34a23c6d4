"""TEST INFRASTRUCTURE: a stack machine for the reference's DEPLOYED verifier bytecode (release-v0.13.1/evm_verifier.bin) -- the
artefact `EVMVerifier::verify_evm_proof` hands to an EVM (/root/reference/integration/tests/e2e_tests.rs:185-199, BASELINE
configs[4] "verifier bytecode check").  The file is creation code: its constructor copies the runtime out of itself and
returns it; `deploy()` runs that, `call()` runs the runtime on a calldata blob.

Only the opcodes the file contains are implemented (arithmetic mod 2^256, ADDMOD / MULMOD, comparisons, bit ops, KECCAK256,
CALLDATALOAD, CODECOPY, memory, JUMP / JUMPI with JUMPDEST analysis, PUSH1..32, DUP, SWAP, GAS, STATICCALL to the precompiles
0x5-0x8, RETURN / REVERT / INVALID); anything else raises, so a silent mis-execution is not possible.  Memory, Keccak and the
precompile plumbing are the ones of tests/yul_verifier.py -- the elliptic-curve precompiles stay pluggable (big-integer model or
the product's host code)."""
from __future__ import annotations

from yul_verifier import M256, Halt, Machine, keccak256


class Evm(Machine):
    def __init__(self, code: bytes, calldata: bytes, ec_add, ec_mul, ec_pairing):
        super().__init__(calldata, ec_add, ec_mul, ec_pairing)
        self.code = code
        self.jumpdests = set()
        i = 0
        while i < len(code):
            op = code[i]
            if op == 0x5B:
                self.jumpdests.add(i)
            i += 1 + (op - 0x5F if 0x60 <= op <= 0x7F else 0)
        self.returndata = b""
        self.steps = 0

    def execute(self):
        """-> True if the code stops with RETURN / STOP, False on REVERT / INVALID / a bad jump"""
        code, st, pc = self.code, [], 0
        push, pop = st.append, st.pop
        while True:
            if pc >= len(code):
                return True  # running off the end is STOP
            op = code[pc]
            self.steps += 1
            pc += 1
            if 0x60 <= op <= 0x7F:  # PUSHn
                n = op - 0x5F
                push(int.from_bytes(code[pc:pc + n].ljust(n, b"\0"), "big"))
                pc += n
            elif 0x80 <= op <= 0x8F:  # DUPn
                push(st[-(op - 0x7F)])
            elif 0x90 <= op <= 0x9F:  # SWAPn
                n = op - 0x8F
                st[-1], st[-1 - n] = st[-1 - n], st[-1]
            elif op == 0x00:
                return True
            elif op == 0x01: push((pop() + pop()) & M256)
            elif op == 0x03: a = pop(); push((a - pop()) & M256)
            elif op == 0x06: a, b = pop(), pop(); push(a % b if b else 0)
            elif op == 0x08: a, b, m = pop(), pop(), pop(); push((a + b) % m if m else 0)
            elif op == 0x09: a, b, m = pop(), pop(), pop(); push((a * b) % m if m else 0)
            elif op == 0x10: a = pop(); push(int(a < pop()))
            elif op == 0x14: push(int(pop() == pop()))
            elif op == 0x15: push(int(pop() == 0))
            elif op == 0x16: push(pop() & pop())
            elif op == 0x17: push(pop() | pop())
            elif op == 0x1B: s = pop(); v = pop(); push((v << s) & M256 if s < 256 else 0)
            elif op == 0x20:
                p, n = pop(), pop()
                self.keccak_calls += 1
                push(int.from_bytes(keccak256(self.mread(p, n)), "big"))
            elif op == 0x35:
                p = pop()
                push(int.from_bytes((self.calldata[p:p + 32] + bytes(32))[:32], "big"))
            elif op == 0x39:  # CODECOPY(dest, offset, size)
                d, o, n = pop(), pop(), pop()
                self._grow(d + n)
                self.mem[d:d + n] = code[o:o + n].ljust(n, b"\0")
            elif op == 0x50: pop()
            elif op == 0x51: push(self.mload(pop()))
            elif op == 0x52: p = pop(); self.mstore(p, pop())
            elif op == 0x53: p = pop(); v = pop(); self._grow(p + 1); self.mem[p] = v & 0xFF
            elif op == 0x56:
                pc = pop()
                if pc not in self.jumpdests:
                    return False
            elif op == 0x57:
                dest, cond = pop(), pop()
                if cond:
                    if dest not in self.jumpdests:
                        return False
                    pc = dest
            elif op == 0x5A: push(M256)
            elif op == 0x5B: pass
            elif op == 0xF3:
                p, n = pop(), pop()
                self.returndata = self.mread(p, n)
                return True
            elif op == 0xFA:  # STATICCALL(gas, addr, in, insize, out, outsize)
                pop()
                addr, ip, isz, opos, osz = pop(), pop(), pop(), pop(), pop()
                push(self.staticcall(addr, ip, isz, opos, osz))
            elif op == 0xFD or op == 0xFE:
                return False
            else:
                raise NotImplementedError(f"evm opcode 0x{op:02x} at {pc - 1}")


def deploy(creation_code: bytes) -> bytes:
    """run the constructor (no calldata, no precompiles) and return the runtime code it hands back"""
    none = lambda *a: (_ for _ in ()).throw(ValueError())
    m = Evm(creation_code, b"", none, none, none)
    if not m.execute() or not m.returndata:
        raise RuntimeError("the creation code did not return a runtime")
    return m.returndata


def call(runtime: bytes, calldata: bytes, ec_add, ec_mul, ec_pairing):
    m = Evm(runtime, calldata, ec_add, ec_mul, ec_pairing)
    try:
        ok = m.execute()
    except Halt as h:  # not raised by Evm itself; kept for symmetry with the Yul machine
        ok = not h.reverted
    return ok, m
