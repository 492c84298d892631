// wasm_interp.cpp — TEST INFRASTRUCTURE.  A small WebAssembly (MVP + sign-ext + sat-trunc + bulk-memory copy/fill)
// interpreter whose only job is to EXECUTE THE REFERENCE ITSELF: the prebuilt ReleaseFast build of resolve2d that the
// reference ships as demos/web/public/resolve2d.wasm (zero imports).  tests/golden/
// make_wasm_golden.py drives it through the module's own exports (solverInit, setup_0_*, solverProcess, the getters
// and setters of src/wasm_root.zig) and records golden vectors that pin the C++ oracle bit for bit.
//
// No reference source is copied or compiled here: the binary is read from the path given on the command line (it
// exists only in the build container, /root/reference; the golden vectors are what travels).
//
// usage: wasm_run <resolve2d.wasm> <scene: 0_1|0_3> <steps> <dt> <sub_steps> <iters> [driven] [remove:<step>:<id>,<id>...] [dump:<s1>,<s2>...]
// output: one JSON object on stdout.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

static void die(const char* m) {
    fprintf(stderr, "wasm_interp: %s\n", m);
    exit(2);
}

struct Reader {
    const u8* p;
    const u8* end;
    u8 byte() {
        if (p >= end) die("eof");
        return *p++;
    }
    u64 uleb() {
        u64 r = 0;
        int s = 0;
        for (;;) {
            u8 b = byte();
            r |= (u64)(b & 0x7f) << s;
            s += 7;
            if (!(b & 0x80)) break;
        }
        return r;
    }
    i64 sleb() {
        i64 r = 0;
        int s = 0;
        u8 b;
        do {
            b = byte();
            r |= (i64)(b & 0x7f) << s;
            s += 7;
        } while (b & 0x80);
        if (s < 64 && (b & 0x40)) r |= -((i64)1 << s);
        return r;
    }
    std::string name() {
        u32 n = (u32)uleb();
        std::string s((const char*)p, n);
        p += n;
        return s;
    }
};

struct FuncType {
    std::vector<u8> params, results;
};
struct Ins {
    uint16_t op;   // 0xFCxx for the prefixed ones
    u32 a = 0;     // depth / index / matching end
    u64 b = 0;     // const bits / memory offset / else position
};
struct Func {
    u32 type = 0;
    std::vector<u8> locals;  // types of the declared locals
    std::vector<Ins> code;
};
struct Label {
    u32 cont;    // pc to continue at when branched to (loop: its first instruction; block/if: after its end)
    u32 height;  // operand stack height at entry
    u32 arity;   // values carried by a branch
    bool is_loop;
};

struct Module {
    std::vector<FuncType> types;
    std::vector<Func> funcs;
    std::vector<u32> table;
    std::vector<u8> mem;
    std::vector<u64> globals;
    std::map<std::string, u32> exports;
    std::vector<std::vector<u32>> br_tables;
    std::vector<u64> stack;
    u64 executed = 0;

    u64 const_expr(Reader& r) {
        u64 v = 0;
        for (;;) {
            u8 op = r.byte();
            if (op == 0x0b) break;
            if (op == 0x41) v = (u32)(i32)r.sleb();
            else if (op == 0x42) v = (u64)r.sleb();
            else if (op == 0x23) v = globals[(u32)r.uleb()];
            else die("unsupported const expr");
        }
        return v;
    }

    void decode_body(Reader& r, Func& f) {
        std::vector<u32> open;  // indices of open block/loop/if
        for (;;) {
            Ins in;
            u8 op = r.byte();
            in.op = op;
            switch (op) {
                case 0x02: case 0x03: case 0x04: {  // block loop if
                    i64 bt = r.sleb();
                    in.b = (bt == -64) ? 0 : 1;  // arity of the block result (void or one value)
                    if (bt >= 0) in.b = types[(u32)bt].results.size();
                    open.push_back((u32)f.code.size());
                } break;
                case 0x05: {  // else
                    f.code[open.back()].b |= (u64)f.code.size() << 32;
                } break;
                case 0x0b: {  // end
                    if (open.empty()) {
                        f.code.push_back(in);
                        return;
                    }
                    f.code[open.back()].a = (u32)f.code.size();
                    open.pop_back();
                } break;
                case 0x0c: case 0x0d: in.a = (u32)r.uleb(); break;  // br br_if
                case 0x0e: {
                    u32 n = (u32)r.uleb();
                    std::vector<u32> t(n + 1);
                    for (u32 k = 0; k <= n; ++k) t[k] = (u32)r.uleb();
                    in.a = (u32)br_tables.size();
                    br_tables.push_back(t);
                } break;
                case 0x10: in.a = (u32)r.uleb(); break;  // call
                case 0x11: in.a = (u32)r.uleb(); r.uleb(); break;  // call_indirect type, table
                case 0x20: case 0x21: case 0x22: case 0x23: case 0x24: in.a = (u32)r.uleb(); break;
                case 0x3f: case 0x40: r.byte(); break;  // memory.size / grow
                case 0x41: in.b = (u32)(i32)r.sleb(); break;
                case 0x42: in.b = (u64)r.sleb(); break;
                case 0x43: { u32 v; memcpy(&v, r.p, 4); r.p += 4; in.b = v; } break;
                case 0x44: { u64 v; memcpy(&v, r.p, 8); r.p += 8; in.b = v; } break;
                case 0xfc: {
                    u32 sub = (u32)r.uleb();
                    in.op = (uint16_t)(0xfc00 | sub);
                    if (sub == 10) { r.byte(); r.byte(); }
                    else if (sub == 11) r.byte();
                    else if (sub > 7) die("unsupported 0xfc op");
                } break;
                default:
                    if (op >= 0x28 && op <= 0x3e) {  // loads / stores: align, offset
                        r.uleb();
                        in.b = r.uleb();
                    }
            }
            f.code.push_back(in);
        }
    }

    void load(const std::vector<u8>& bin) {
        Reader r{bin.data(), bin.data() + bin.size()};
        if (bin.size() < 8 || memcmp(bin.data(), "\0asm", 4)) die("not a wasm file");
        r.p += 8;
        std::vector<u32> func_types;
        while (r.p < r.end) {
            u8 id = r.byte();
            u32 size = (u32)r.uleb();
            Reader s{r.p, r.p + size};
            r.p += size;
            switch (id) {
                case 1: {
                    u32 n = (u32)s.uleb();
                    for (u32 k = 0; k < n; ++k) {
                        if (s.byte() != 0x60) die("bad functype");
                        FuncType t;
                        u32 np = (u32)s.uleb();
                        for (u32 q = 0; q < np; ++q) t.params.push_back(s.byte());
                        u32 nr = (u32)s.uleb();
                        for (u32 q = 0; q < nr; ++q) t.results.push_back(s.byte());
                        types.push_back(t);
                    }
                } break;
                case 2: if (s.uleb() != 0) die("module has imports"); break;
                case 3: {
                    u32 n = (u32)s.uleb();
                    for (u32 k = 0; k < n; ++k) func_types.push_back((u32)s.uleb());
                } break;
                case 4: {
                    u32 n = (u32)s.uleb();
                    for (u32 k = 0; k < n; ++k) {
                        s.byte();
                        u8 flag = s.byte();
                        u32 mn = (u32)s.uleb();
                        if (flag & 1) s.uleb();
                        table.assign(mn, 0xffffffffu);
                    }
                } break;
                case 5: {
                    u32 n = (u32)s.uleb();
                    for (u32 k = 0; k < n; ++k) {
                        u8 flag = s.byte();
                        u32 mn = (u32)s.uleb();
                        if (flag & 1) s.uleb();
                        mem.assign((size_t)mn * 65536, 0);
                    }
                } break;
                case 6: {
                    u32 n = (u32)s.uleb();
                    for (u32 k = 0; k < n; ++k) {
                        s.byte();
                        s.byte();
                        globals.push_back(const_expr(s));
                    }
                } break;
                case 7: {
                    u32 n = (u32)s.uleb();
                    for (u32 k = 0; k < n; ++k) {
                        std::string nm = s.name();
                        u8 kind = s.byte();
                        u32 idx = (u32)s.uleb();
                        if (kind == 0) exports[nm] = idx;
                    }
                } break;
                case 9: {
                    u32 n = (u32)s.uleb();
                    for (u32 k = 0; k < n; ++k) {
                        u32 flag = (u32)s.uleb();
                        if (flag != 0) die("unsupported element segment");
                        u32 off = (u32)const_expr(s);
                        u32 cnt = (u32)s.uleb();
                        if (table.size() < off + cnt) table.resize(off + cnt, 0xffffffffu);
                        for (u32 q = 0; q < cnt; ++q) table[off + q] = (u32)s.uleb();
                    }
                } break;
                case 10: {
                    u32 n = (u32)s.uleb();
                    funcs.resize(n);
                    for (u32 k = 0; k < n; ++k) {
                        u32 sz = (u32)s.uleb();
                        Reader b{s.p, s.p + sz};
                        s.p += sz;
                        Func& f = funcs[k];
                        f.type = func_types[k];
                        u32 groups = (u32)b.uleb();
                        for (u32 g = 0; g < groups; ++g) {
                            u32 cnt = (u32)b.uleb();
                            u8 t = b.byte();
                            f.locals.insert(f.locals.end(), cnt, t);
                        }
                        decode_body(b, f);
                    }
                } break;
                case 11: {
                    u32 n = (u32)s.uleb();
                    for (u32 k = 0; k < n; ++k) {
                        u32 flag = (u32)s.uleb();
                        if (flag == 1) { u32 len = (u32)s.uleb(); s.p += len; continue; }
                        if (flag == 2) s.uleb();
                        u32 off = (u32)const_expr(s);
                        u32 len = (u32)s.uleb();
                        if (mem.size() < (size_t)off + len) die("data segment out of range");
                        memcpy(&mem[off], s.p, len);
                        s.p += len;
                    }
                } break;
                default: break;  // custom, data count, start (none)
            }
        }
    }

    // ---- execution -----------------------------------------------------------------------------------------------
    static float f32_of(u64 v) { u32 b = (u32)v; float f; memcpy(&f, &b, 4); return f; }
    static u64 of_f32(float f) { u32 b; memcpy(&b, &f, 4); return b; }
    static double f64_of(u64 v) { double d; memcpy(&d, &v, 8); return d; }
    static u64 of_f64(double d) { u64 b; memcpy(&b, &d, 8); return b; }
    u64 pop() { u64 v = stack.back(); stack.pop_back(); return v; }
    void push(u64 v) { stack.push_back(v); }
    template <class T> T ld(u64 addr) {
        if (addr + sizeof(T) > mem.size()) die("load out of bounds");
        T v; memcpy(&v, &mem[addr], sizeof(T)); return v;
    }
    template <class T> void st(u64 addr, T v) {
        if (addr + sizeof(T) > mem.size()) die("store out of bounds");
        memcpy(&mem[addr], &v, sizeof(T));
    }
    template <class I, class F> static I trunc_sat(F f, I lo, I hi) {
        if (f != f) return 0;
        if (f <= (F)lo) return lo;
        if (f >= (F)hi) return hi;
        return (I)f;
    }

    void call(u32 fi) {
        const Func& f = funcs[fi];
        const FuncType& ft = types[f.type];
        std::vector<u64> locals(ft.params.size() + f.locals.size(), 0);
        for (size_t k = ft.params.size(); k-- > 0;) locals[k] = pop();
        std::vector<Label> labels;
        labels.push_back({(u32)f.code.size(), (u32)stack.size(), (u32)ft.results.size(), false});
        const Ins* code = f.code.data();
        u32 pc = 0;
        for (;;) {
            if (pc >= f.code.size()) break;
            const Ins& in = code[pc++];
            executed++;
            switch (in.op) {
                case 0x00: die("unreachable executed");
                case 0x01: break;
                case 0x02: labels.push_back({in.a + 1, (u32)stack.size(), (u32)(in.b & 0xffffffffu), false}); break;  // block: continue after end
                case 0x03: labels.push_back({pc, (u32)stack.size(), 0, true}); break;                                 // loop: continue at its first instruction
                case 0x04: {
                    u32 c = (u32)pop();
                    u32 els = (u32)(in.b >> 32);
                    labels.push_back({in.a + 1, (u32)stack.size(), (u32)(in.b & 0xffffffffu), false});
                    if (!c) {
                        if (els) pc = els + 1;
                        else { pc = in.a + 1; labels.pop_back(); }
                    }
                } break;
                case 0x05: {  // reached the else from the then-branch: skip to end
                    const Label l = labels.back();
                    labels.pop_back();
                    pc = l.cont;
                } break;
                case 0x0b:
                    if (labels.size() > 1) labels.pop_back();
                    else goto done;
                    break;
                case 0x0c: case 0x0d: case 0x0e: {
                    u32 depth;
                    if (in.op == 0x0c) depth = in.a;
                    else if (in.op == 0x0d) { if (!(u32)pop()) break; depth = in.a; }
                    else { u32 i = (u32)pop(); const auto& t = br_tables[in.a]; depth = i < t.size() - 1 ? t[i] : t.back(); }
                    if (depth >= labels.size()) die("bad branch depth");
                    const size_t li = labels.size() - 1 - depth;
                    const Label l = labels[li];
                    if (l.arity) {
                        u64 v = stack.back();
                        stack.resize(l.height);
                        stack.push_back(v);
                    } else {
                        stack.resize(l.height);
                    }
                    if (li == 0) goto done;  // branch to the function label = return
                    if (l.is_loop) {
                        labels.resize(li + 1);  // the loop keeps its label
                    } else {
                        labels.resize(li);
                    }
                    pc = l.cont;
                } break;
                case 0x0f: {
                    const Label l = labels[0];
                    if (l.arity) { u64 v = stack.back(); stack.resize(l.height); stack.push_back(v); }
                    else stack.resize(l.height);
                    goto done;
                }
                case 0x10: call(in.a); break;
                case 0x11: {
                    u32 i = (u32)pop();
                    if (i >= table.size() || table[i] == 0xffffffffu) die("bad indirect call");
                    call(table[i]);
                } break;
                case 0x1a: pop(); break;
                case 0x1b: { u32 c = (u32)pop(); u64 b = pop(), a = pop(); push(c ? a : b); } break;
                case 0x20: push(locals[in.a]); break;
                case 0x21: locals[in.a] = pop(); break;
                case 0x22: locals[in.a] = stack.back(); break;
                case 0x23: push(globals[in.a]); break;
                case 0x24: globals[in.a] = pop(); break;
                // loads
                case 0x28: push(ld<u32>((u32)pop() + in.b)); break;
                case 0x29: push(ld<u64>((u32)pop() + in.b)); break;
                case 0x2a: push(ld<u32>((u32)pop() + in.b)); break;
                case 0x2b: push(ld<u64>((u32)pop() + in.b)); break;
                case 0x2c: push((u32)(i32)ld<int8_t>((u32)pop() + in.b)); break;
                case 0x2d: push(ld<u8>((u32)pop() + in.b)); break;
                case 0x2e: push((u32)(i32)ld<int16_t>((u32)pop() + in.b)); break;
                case 0x2f: push(ld<uint16_t>((u32)pop() + in.b)); break;
                case 0x30: push((u64)(i64)ld<int8_t>((u32)pop() + in.b)); break;
                case 0x31: push(ld<u8>((u32)pop() + in.b)); break;
                case 0x32: push((u64)(i64)ld<int16_t>((u32)pop() + in.b)); break;
                case 0x33: push(ld<uint16_t>((u32)pop() + in.b)); break;
                case 0x34: push((u64)(i64)ld<i32>((u32)pop() + in.b)); break;
                case 0x35: push(ld<u32>((u32)pop() + in.b)); break;
                // stores
                case 0x36: case 0x38: { u32 v = (u32)pop(); st<u32>((u32)pop() + in.b, v); } break;
                case 0x37: case 0x39: { u64 v = pop(); st<u64>((u32)pop() + in.b, v); } break;
                case 0x3a: case 0x3c: { u8 v = (u8)pop(); st<u8>((u32)pop() + in.b, v); } break;
                case 0x3b: case 0x3d: { uint16_t v = (uint16_t)pop(); st<uint16_t>((u32)pop() + in.b, v); } break;
                case 0x3e: { u32 v = (u32)pop(); st<u32>((u32)pop() + in.b, v); } break;
                case 0x3f: push((u32)(mem.size() / 65536)); break;
                case 0x40: {
                    u32 n = (u32)pop();
                    u32 old = (u32)(mem.size() / 65536);
                    if ((u64)old + n > 32768) push(0xffffffffu);
                    else { mem.resize((size_t)(old + n) * 65536, 0); push(old); }
                } break;
                case 0x41: case 0x42: case 0x43: case 0x44: push(in.b); break;
#define CMP32(OP, T) { T b = (T)(u32)pop(), a = (T)(u32)pop(); push(a OP b ? 1 : 0); } break
#define CMP64(OP, T) { T b = (T)pop(), a = (T)pop(); push(a OP b ? 1 : 0); } break
                case 0x45: push((u32)pop() == 0); break;
                case 0x46: CMP32(==, u32); case 0x47: CMP32(!=, u32);
                case 0x48: CMP32(<, i32); case 0x49: CMP32(<, u32); case 0x4a: CMP32(>, i32); case 0x4b: CMP32(>, u32);
                case 0x4c: CMP32(<=, i32); case 0x4d: CMP32(<=, u32); case 0x4e: CMP32(>=, i32); case 0x4f: CMP32(>=, u32);
                case 0x50: push(pop() == 0); break;
                case 0x51: CMP64(==, u64); case 0x52: CMP64(!=, u64);
                case 0x53: CMP64(<, i64); case 0x54: CMP64(<, u64); case 0x55: CMP64(>, i64); case 0x56: CMP64(>, u64);
                case 0x57: CMP64(<=, i64); case 0x58: CMP64(<=, u64); case 0x59: CMP64(>=, i64); case 0x5a: CMP64(>=, u64);
#define FCMP32(OP) { float b = f32_of(pop()), a = f32_of(pop()); push(a OP b ? 1 : 0); } break
#define FCMP64(OP) { double b = f64_of(pop()), a = f64_of(pop()); push(a OP b ? 1 : 0); } break
                case 0x5b: FCMP32(==); case 0x5c: FCMP32(!=); case 0x5d: FCMP32(<); case 0x5e: FCMP32(>); case 0x5f: FCMP32(<=); case 0x60: FCMP32(>=);
                case 0x61: FCMP64(==); case 0x62: FCMP64(!=); case 0x63: FCMP64(<); case 0x64: FCMP64(>); case 0x65: FCMP64(<=); case 0x66: FCMP64(>=);
                case 0x67: { u32 a = (u32)pop(); push(a ? (u32)__builtin_clz(a) : 32); } break;
                case 0x68: { u32 a = (u32)pop(); push(a ? (u32)__builtin_ctz(a) : 32); } break;
                case 0x69: push((u32)__builtin_popcount((u32)pop())); break;
#define BIN32(EXPR) { u32 b = (u32)pop(), a = (u32)pop(); (void)a; (void)b; push((u32)(EXPR)); } break
                case 0x6a: BIN32(a + b); case 0x6b: BIN32(a - b); case 0x6c: BIN32(a * b);
                case 0x6d: { i32 b = (i32)pop(), a = (i32)pop(); if (!b) die("div by zero"); push((u32)((a == INT32_MIN && b == -1) ? a : a / b)); } break;
                case 0x6e: { u32 b = (u32)pop(), a = (u32)pop(); if (!b) die("div by zero"); push(a / b); } break;
                case 0x6f: { i32 b = (i32)pop(), a = (i32)pop(); if (!b) die("div by zero"); push((u32)((b == -1) ? 0 : a % b)); } break;
                case 0x70: { u32 b = (u32)pop(), a = (u32)pop(); if (!b) die("div by zero"); push(a % b); } break;
                case 0x71: BIN32(a & b); case 0x72: BIN32(a | b); case 0x73: BIN32(a ^ b);
                case 0x74: BIN32(a << (b & 31)); case 0x75: BIN32((u32)((i32)a >> (b & 31))); case 0x76: BIN32(a >> (b & 31));
                case 0x77: BIN32((a << (b & 31)) | (a >> ((32 - (b & 31)) & 31)));
                case 0x78: BIN32((a >> (b & 31)) | (a << ((32 - (b & 31)) & 31)));
                case 0x79: { u64 a = pop(); push(a ? (u64)__builtin_clzll(a) : 64); } break;
                case 0x7a: { u64 a = pop(); push(a ? (u64)__builtin_ctzll(a) : 64); } break;
                case 0x7b: push((u64)__builtin_popcountll(pop())); break;
#define BIN64(EXPR) { u64 b = pop(), a = pop(); (void)a; (void)b; push((u64)(EXPR)); } break
                case 0x7c: BIN64(a + b); case 0x7d: BIN64(a - b); case 0x7e: BIN64(a * b);
                case 0x7f: { i64 b = (i64)pop(), a = (i64)pop(); if (!b) die("div by zero"); push((u64)((a == INT64_MIN && b == -1) ? a : a / b)); } break;
                case 0x80: { u64 b = pop(), a = pop(); if (!b) die("div by zero"); push(a / b); } break;
                case 0x81: { i64 b = (i64)pop(), a = (i64)pop(); if (!b) die("div by zero"); push((u64)((b == -1) ? 0 : a % b)); } break;
                case 0x82: { u64 b = pop(), a = pop(); if (!b) die("div by zero"); push(a % b); } break;
                case 0x83: BIN64(a & b); case 0x84: BIN64(a | b); case 0x85: BIN64(a ^ b);
                case 0x86: BIN64(a << (b & 63)); case 0x87: BIN64((u64)((i64)a >> (b & 63))); case 0x88: BIN64(a >> (b & 63));
                case 0x89: BIN64((a << (b & 63)) | (a >> ((64 - (b & 63)) & 63)));
                case 0x8a: BIN64((a >> (b & 63)) | (a << ((64 - (b & 63)) & 63)));
                // f32
                case 0x8b: push(of_f32(fabsf(f32_of(pop())))); break;
                case 0x8c: push((u32)pop() ^ 0x80000000u); break;
                case 0x8d: push(of_f32(ceilf(f32_of(pop())))); break;
                case 0x8e: push(of_f32(floorf(f32_of(pop())))); break;
                case 0x8f: push(of_f32(truncf(f32_of(pop())))); break;
                case 0x90: push(of_f32(nearbyintf(f32_of(pop())))); break;
                case 0x91: push(of_f32(sqrtf(f32_of(pop())))); break;
#define FBIN32(EXPR) { float b = f32_of(pop()), a = f32_of(pop()); push(of_f32(EXPR)); } break
                case 0x92: FBIN32(a + b); case 0x93: FBIN32(a - b); case 0x94: FBIN32(a * b); case 0x95: FBIN32(a / b);
                case 0x96: FBIN32((a != a || b != b) ? NAN : fminf(a, b)); case 0x97: FBIN32((a != a || b != b) ? NAN : fmaxf(a, b));
                case 0x98: FBIN32(copysignf(a, b));
                // f64
                case 0x99: push(of_f64(fabs(f64_of(pop())))); break;
                case 0x9a: push(pop() ^ 0x8000000000000000ull); break;
                case 0x9b: push(of_f64(ceil(f64_of(pop())))); break;
                case 0x9c: push(of_f64(floor(f64_of(pop())))); break;
                case 0x9d: push(of_f64(trunc(f64_of(pop())))); break;
                case 0x9e: push(of_f64(nearbyint(f64_of(pop())))); break;
                case 0x9f: push(of_f64(sqrt(f64_of(pop())))); break;
#define FBIN64(EXPR) { double b = f64_of(pop()), a = f64_of(pop()); push(of_f64(EXPR)); } break
                case 0xa0: FBIN64(a + b); case 0xa1: FBIN64(a - b); case 0xa2: FBIN64(a * b); case 0xa3: FBIN64(a / b);
                case 0xa4: FBIN64((a != a || b != b) ? NAN : fmin(a, b)); case 0xa5: FBIN64((a != a || b != b) ? NAN : fmax(a, b));
                case 0xa6: FBIN64(copysign(a, b));
                // conversions
                case 0xa7: push((u32)pop()); break;
                case 0xa8: push((u32)(i32)f32_of(pop())); break;
                case 0xa9: push((u32)f32_of(pop())); break;
                case 0xaa: push((u32)(i32)f64_of(pop())); break;
                case 0xab: push((u32)f64_of(pop())); break;
                case 0xac: push((u64)(i64)(i32)(u32)pop()); break;
                case 0xad: push((u64)(u32)pop()); break;
                case 0xae: push((u64)(i64)f32_of(pop())); break;
                case 0xaf: push((u64)f32_of(pop())); break;
                case 0xb0: push((u64)(i64)f64_of(pop())); break;
                case 0xb1: push((u64)f64_of(pop())); break;
                case 0xb2: push(of_f32((float)(i32)(u32)pop())); break;
                case 0xb3: push(of_f32((float)(u32)pop())); break;
                case 0xb4: push(of_f32((float)(i64)pop())); break;
                case 0xb5: push(of_f32((float)pop())); break;
                case 0xb6: push(of_f32((float)f64_of(pop()))); break;
                case 0xb7: push(of_f64((double)(i32)(u32)pop())); break;
                case 0xb8: push(of_f64((double)(u32)pop())); break;
                case 0xb9: push(of_f64((double)(i64)pop())); break;
                case 0xba: push(of_f64((double)pop())); break;
                case 0xbb: push(of_f64((double)f32_of(pop()))); break;
                case 0xbc: case 0xbd: case 0xbe: case 0xbf: break;  // reinterprets: bit patterns are stored as is
                case 0xc0: push((u32)(i32)(int8_t)pop()); break;
                case 0xc1: push((u32)(i32)(int16_t)pop()); break;
                case 0xc2: push((u64)(i64)(int8_t)pop()); break;
                case 0xc3: push((u64)(i64)(int16_t)pop()); break;
                case 0xc4: push((u64)(i64)(i32)pop()); break;
                case 0xfc00: push((u32)trunc_sat<i32, float>(f32_of(pop()), INT32_MIN, INT32_MAX)); break;
                case 0xfc01: push(trunc_sat<u32, float>(f32_of(pop()), 0u, UINT32_MAX)); break;
                case 0xfc02: push((u32)trunc_sat<i32, double>(f64_of(pop()), INT32_MIN, INT32_MAX)); break;
                case 0xfc03: push(trunc_sat<u32, double>(f64_of(pop()), 0u, UINT32_MAX)); break;
                case 0xfc04: push((u64)trunc_sat<i64, float>(f32_of(pop()), INT64_MIN, INT64_MAX)); break;
                case 0xfc05: push(trunc_sat<u64, float>(f32_of(pop()), 0ull, UINT64_MAX)); break;
                case 0xfc06: push((u64)trunc_sat<i64, double>(f64_of(pop()), INT64_MIN, INT64_MAX)); break;
                case 0xfc07: push(trunc_sat<u64, double>(f64_of(pop()), 0ull, UINT64_MAX)); break;
                case 0xfc0a: { u32 n = (u32)pop(), s = (u32)pop(), d = (u32)pop(); if ((u64)s + n > mem.size() || (u64)d + n > mem.size()) die("memory.copy oob"); memmove(&mem[d], &mem[s], n); } break;
                case 0xfc0b: { u32 n = (u32)pop(), v = (u32)pop(), d = (u32)pop(); if ((u64)d + n > mem.size()) die("memory.fill oob"); memset(&mem[d], (int)v, n); } break;
                default: fprintf(stderr, "opcode 0x%x\n", in.op); die("unsupported opcode");
            }
        }
    done:
        return;
    }

    u64 invoke(const std::string& name, std::vector<u64> args) {
        auto it = exports.find(name);
        if (it == exports.end()) { fprintf(stderr, "%s\n", name.c_str()); die("no such export"); }
        const FuncType& ft = types[funcs[it->second].type];
        stack.clear();
        for (u64 a : args) push(a);
        call(it->second);
        return ft.results.empty() ? 0 : stack.back();
    }
};

// ---- driver ---------------------------------------------------------------------------------------------------------
static u64 fnv_words(u64 h, const u32* w, size_t n) {
    for (size_t k = 0; k < n; ++k)
        for (int b = 0; b < 4; ++b) {
            h ^= (w[k] >> (8 * b)) & 0xffu;
            h *= 0x100000001B3ull;
        }
    return h;
}
static u64 f32bits(float f) { u32 b; memcpy(&b, &f, 4); return b; }

int main(int argc, char** argv) {
    if (argc < 7) die("usage: wasm_run <wasm> <0_1|0_3> <steps> <dt> <sub_steps> <iters> [driven] [remove:<step>:<id>,...] [dump:<s>,...]");
    FILE* fp = fopen(argv[1], "rb");
    if (!fp) die("cannot open wasm");
    std::vector<u8> bin;
    u8 buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, fp)) > 0) bin.insert(bin.end(), buf, buf + n);
    fclose(fp);
    Module m;
    m.load(bin);
    const std::string scene = argv[2];
    const int steps = atoi(argv[3]);
    const float dt = (float)(1.0 / atof(argv[4]));  // argument is the rate: dt = 1/rate evaluated in f32 like the tests (1.0f/60.0f)
    const float dtf = 1.0f / (float)atof(argv[4]);
    (void)dt;
    const u32 S = (u32)atoi(argv[5]), I = (u32)atoi(argv[6]);
    bool driven = false;
    std::map<int, std::vector<u32>> removals;
    std::vector<int> dumps;
    for (int a = 7; a < argc; ++a) {
        std::string s = argv[a];
        if (s == "driven") driven = true;
        else if (s.rfind("remove:", 0) == 0) {
            size_t c = s.find(':', 7);
            int step = atoi(s.substr(7, c - 7).c_str());
            std::string ids = s.substr(c + 1);
            size_t pos = 0;
            while (pos < ids.size()) {
                size_t e = ids.find(',', pos);
                if (e == std::string::npos) e = ids.size();
                removals[step].push_back((u32)atoi(ids.substr(pos, e - pos).c_str()));
                pos = e + 1;
            }
        } else if (s.rfind("dump:", 0) == 0) {
            std::string l = s.substr(5);
            size_t pos = 0;
            while (pos < l.size()) {
                size_t e = l.find(',', pos);
                if (e == std::string::npos) e = l.size();
                dumps.push_back(atoi(l.substr(pos, e - pos).c_str()));
                pos = e + 1;
            }
        }
    }
    if (!m.invoke("solverInit", {f32bits(2.0f), 4})) die("solverInit failed");
    if (!m.invoke(scene == "0_1" ? "setup_0_1_car_platformer" : "setup_0_3_many_boxes", {})) die("setup failed");
    auto ptr_of = [&](u32 id) { return (u64)(u32)m.invoke("getRigidBodyPtrFromId", {id}); };
    printf("{\"scene\": \"%s\", \"dt_rate\": %s, \"sub_steps\": %u, \"iters\": %u, \"driven\": %s, \"steps\": [\n", scene.c_str(), argv[4], S, I,
           driven ? "true" : "false");
    for (int step = 0; step <= steps; ++step) {
        if (step > 0) {
            if (removals.count(step))
                for (u32 id : removals[step])
                    if (!m.invoke("solverRemoveBodyById", {id})) die("remove failed");
            if (driven) {  // SURVEY F.3 / demos/native/src/main.zig:124-133
                m.invoke("setRigidBodyAngularMomentum", {ptr_of(4), f32bits(-20.0f)});
                m.invoke("setRigidBodyAngularMomentum", {ptr_of(5), f32bits(-20.0f)});
                m.invoke("setRigidBodyTorque", {ptr_of(3), f32bits(400.0f)});
                m.invoke("setRigidBodyTorque", {ptr_of(109), f32bits(50.0f)});
                m.invoke("setRigidBodyForceX", {ptr_of(3), f32bits(3.0f)});
            }
            if (!m.invoke("solverProcess", {f32bits(dtf), S, I})) die("solverProcess failed");
        }
        const u32 nb = (u32)m.invoke("solverGetNumBodies", {});
        u64 hs = 0x14650FB0739D0383ull, ha = 0x14650FB0739D0383ull;
        bool dump = false;
        for (int d : dumps) dump |= (d == step);
        std::string raw;
        for (u32 i = 0; i < nb; ++i) {
            const u32 id = (u32)m.invoke("solverGetBodyIdBasedOnIter", {i}) & 0xffffu;
            const u64 p = ptr_of(id);
            u32 w[11];
            w[0] = id;
            const char* names[10] = {"getRigidBodyPosX", "getRigidBodyPosY", "getRigidBodyAngle", "getRigidBodyMomentumX", "getRigidBodyMomentumY",
                                     "getRigidBodyAngularMomentum", "getRigidBodyAABBPosX", "getRigidBodyAABBPosY", "getRigidBodyAABBHalfWidth",
                                     "getRigidBodyAABBHalfHeight"};
            for (int k = 0; k < 10; ++k) w[1 + k] = (u32)m.invoke(names[k], {p});
            hs = fnv_words(hs, w, 7);
            u32 wa[5] = {id, w[7], w[8], w[9], w[10]};
            ha = fnv_words(ha, wa, 5);
            if (dump) {
                char line[256];
                snprintf(line, sizeof line, "%s[%u,%u,%u,%u,%u,%u,%u,%u,%u,%u,%u]", i ? "," : "", w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8], w[9], w[10]);
                raw += line;
            }
        }
        printf(" {\"step\": %d, \"n\": %u, \"state\": \"%016llx\", \"aabb\": \"%016llx\"%s%s%s}%s\n", step, nb, (unsigned long long)hs,
               (unsigned long long)ha, dump ? ", \"bodies\": [" : "", raw.c_str(), dump ? "]" : "", step == steps ? "" : ",");
    }
    printf("], \"wasm_instructions\": %llu}\n", (unsigned long long)m.executed);
    return 0;
}
