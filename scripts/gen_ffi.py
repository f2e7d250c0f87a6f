#!/usr/bin/env python
"""Generate the jax.ffi registration layer from include/dogstep.h.

    python scripts/gen_ffi.py          # writes exploring-muzero-on-dog_b200/csrc/ffi/dogstep_ffi.cc and
                                       #        exploring-muzero-on-dog_b200/ffi_table.json

One XLA FFI handler per entry point that takes a stream (the host-only utilities — version, last_error, num_actions,
is_sparse, host_split, host_key_chain — are not device calls).  The mapping is mechanical, so a maintainer never writes a
handler by hand and the CPU test (tests/test_ffi.py) fails when the header and the generated files drift apart:

  void* stream                         -> Ctx<PlatformStream<cudaStream_t>>
  const dogstep_*_cfg* (scalars only)  -> one Attr per field                     (the static fields of the reference pytrees)
  struct of leaves (state, tree, replay arrays / batch)
                                       -> per pointer field one operand AND one aliased result (input_output_aliases on the
                                          JAX side: the kernels update leaves in place); per scalar field one Attr
  const T* host_* / agent_type         -> Attrs (host values: the two words of a key, the four agent types)
  const T* p                           -> operand (an empty buffer = NULL)
  T* p                                 -> operand + aliased result (in/out; a pure output's operand is an uninitialised buffer)
  int64_t / int32_t / uint32_t / float -> Attr
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dogstep.h")
OUT_CC = os.path.join(ROOT, "exploring-muzero-on-dog_b200", "csrc", "ffi", "dogstep_ffi.cc")
OUT_JSON = os.path.join(ROOT, "exploring-muzero-on-dog_b200", "ffi_table.json")

HOST_ARRAYS = {"host_rng_key": ("uint32_t", 2), "host_key": ("uint32_t", 2), "agent_type": ("int32_t", 4)}
SCALARS = {"int64_t": "int64_t", "int32_t": "int32_t", "uint32_t": "uint32_t", "float": "float", "int": "int32_t"}


def parse_header(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    structs, aliases = {}, {}
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            first, *rest = [d.strip() for d in decl.split(",")]
            mm = re.match(r"(.*?)(\**)\s*(\w+)$", first)
            base, stars, name = mm.group(1).strip(), mm.group(2), mm.group(3)
            fields.append((base + stars, name))
            for r in rest:
                mm = re.match(r"(\**)\s*(\w+)$", r)
                fields.append((base + mm.group(1), mm.group(2)))
        structs[m.group(2)] = fields
    for m in re.finditer(r"typedef\s+(dogstep_\w+)\s+(dogstep_\w+)\s*;", text):
        aliases[m.group(2)] = m.group(1)
    protos = []
    for m in re.finditer(r"\bint\s+(dogstep_\w+)\s*\(([^)]*)\)\s*;", text):
        params = []
        for p in m.group(2).split(","):
            p = " ".join(p.split())
            if p in ("void", ""):
                continue
            mm = re.match(r"(.*?)(\**)\s*(\w+)$", p)
            params.append(((mm.group(1).strip() + mm.group(2)).replace(" *", "*"), mm.group(3)))
        protos.append((m.group(1), params))
    return structs, aliases, protos


def plan(structs, aliases, protos):
    """-> list of handler descriptions"""
    def resolve(t):
        return aliases.get(t, t)

    handlers = []
    for name, params in protos:
        if not any(t == "void*" and n == "stream" for t, n in params):
            continue
        h = {"name": name, "symbol": "k_" + name, "operands": [], "attrs": [], "call": []}
        for t, n in params:
            base = t.replace("const ", "").rstrip("*").strip()
            is_ptr = t.endswith("*")
            const = t.startswith("const ")
            if t == "void*" and n == "stream":
                h["call"].append(("stream",))
            elif is_ptr and resolve(base) in structs:
                fields = structs[resolve(base)]
                if all("*" not in ft for ft, _ in fields):       # cfg: scalars only
                    for ft, fn in fields:
                        h["attrs"].append({"name": f"{n}_{fn}", "type": SCALARS[ft]})
                    h["call"].append(("cfg", resolve(base), n, [fn for _, fn in fields]))
                else:
                    items = []
                    for ft, fn in fields:
                        if "*" in ft:
                            h["operands"].append({"name": f"{n}_{fn}", "mutable": True, "ctype": ft})
                            items.append(("buf", f"{n}_{fn}", ft))
                        else:
                            h["attrs"].append({"name": f"{n}_{fn}", "type": SCALARS[ft]})
                            items.append(("attr", f"{n}_{fn}", ft))
                    h["call"].append(("struct", resolve(base), n, items))
            elif is_ptr and n in HOST_ARRAYS:
                ct, cnt = HOST_ARRAYS[n]
                for i in range(cnt):
                    h["attrs"].append({"name": f"{n}_{i}", "type": SCALARS[ct]})
                h["call"].append(("host_array", ct, n, cnt))
            elif is_ptr:
                h["operands"].append({"name": n, "mutable": not const, "ctype": t})
                h["call"].append(("ptr", n, t))
            else:
                h["attrs"].append({"name": n, "type": SCALARS[base]})
                h["call"].append(("scalar", n, SCALARS[base]))
        handlers.append(h)
    return handlers


def emit_cc(handlers):
    o = []
    w = o.append
    w("// dogstep_ffi.cc — GENERATED by scripts/gen_ffi.py from include/dogstep.h.  Do not edit.")
    w("//")
    w("// XLA FFI handlers (jax.ffi custom calls) for every stream-taking entry point of libdogstep.so, so that the reference's")
    w("// jitted code (MuZero_det_MADN/game_agent.py:66-84,114, evaluate_agent.py:331-350, the mctx callbacks, vec_replay_buffer.py)")
    w("// can call the CUDA path without leaving the XLA program.  Build where jaxlib is installed:")
    w("//   g++ -O2 -std=c++17 -shared -fPIC -I$(python -c 'import jax.ffi; print(jax.ffi.include_dir())') -Iinclude \\")
    w("//       -I/usr/local/cuda/include exploring-muzero-on-dog_b200/csrc/ffi/dogstep_ffi.cc \\")
    w("//       -Lexploring-muzero-on-dog_b200 -ldogstep -L/usr/local/cuda/lib64 -lcudart -o libdogstep_ffi.so")
    w("// Conventions: see scripts/gen_ffi.py.  Mutable buffers arrive as operand + result; XLA aliases them when the caller passes")
    w("// input_output_aliases (jax_plugin.py does), otherwise the operand is copied into the result first, on the call's stream.")
    w("#if __has_include(\"xla/ffi/api/ffi.h\")")
    w("#include <cstdint>")
    w("#include <string>")
    w("#include <cuda_runtime.h>")
    w("#include \"xla/ffi/api/ffi.h\"")
    w("#include \"dogstep.h\"")
    w("")
    w("namespace ffi = xla::ffi;")
    w("")
    w("namespace {")
    w("// device pointer of an operand (an empty buffer stands for NULL: optional leaves)")
    w("template <typename T>")
    w("inline T* in_ptr(ffi::AnyBuffer b) { return b.element_count() == 0 ? nullptr : reinterpret_cast<T*>(b.untyped_data()); }")
    w("// device pointer of an in/out buffer: the result, holding the operand's bytes (aliased, or copied here)")
    w("template <typename T>")
    w("inline T* io_ptr(ffi::AnyBuffer in, ffi::Result<ffi::AnyBuffer>& out, cudaStream_t stream) {")
    w("  if (out->element_count() == 0) return nullptr;")
    w("  if (in.element_count() != 0 && in.untyped_data() != out->untyped_data())")
    w("    cudaMemcpyAsync(out->untyped_data(), in.untyped_data(), in.size_bytes(), cudaMemcpyDeviceToDevice, stream);")
    w("  return reinterpret_cast<T*>(out->untyped_data());")
    w("}")
    w("inline ffi::Error status(int rc, const char* what) {")
    w("  if (rc == DOGSTEP_OK) return ffi::Error::Success();")
    w("  std::string msg = std::string(what) + (rc == DOGSTEP_ERR_INVALID_ARG ? \": invalid argument\" : rc == DOGSTEP_ERR_UNSUPPORTED")
    w("                        ? \": unsupported configuration\" : std::string(\": CUDA failure: \") + dogstep_last_error());")
    w("  return rc == DOGSTEP_ERR_CUDA ? ffi::Error::Internal(msg) : ffi::Error::InvalidArgument(msg);")
    w("}")
    w("}  // namespace")
    w("")
    for h in handlers:
        ops, attrs = h["operands"], h["attrs"]
        muts = [x for x in ops if x["mutable"]]
        sig = ["cudaStream_t stream"] + [f"ffi::AnyBuffer {x['name']}" for x in ops] + \
              [f"ffi::Result<ffi::AnyBuffer> {x['name']}_out" for x in muts] + [f"{a['type']} {a['name']}" for a in attrs]
        w(f"static ffi::Error {h['name']}_impl({', '.join(sig)}) {{")
        args = []
        mut_names = {x["name"] for x in muts}
        for c in h["call"]:
            if c[0] == "stream":
                args.append("stream")
            elif c[0] == "cfg":
                _, st, n, fields = c
                w(f"  {st} {n}{{{', '.join(f'{n}_{f}' for f in fields)}}};")
                args.append("&" + n)
            elif c[0] == "struct":
                _, st, n, items = c
                inits = []
                for kind, nm, ft in items:
                    if kind == "attr":
                        inits.append(nm)
                    else:
                        base = ft.rstrip("*").strip()
                        inits.append(f"io_ptr<{base}>({nm}, {nm}_out, stream)")
                w(f"  {st} {n}{{{', '.join(inits)}}};")
                args.append("&" + n)
            elif c[0] == "host_array":
                _, ct, n, cnt = c
                w(f"  const {ct} {n}[{cnt}] = {{{', '.join(f'{n}_{i}' for i in range(cnt))}}};")
                args.append(n)
            elif c[0] == "ptr":
                _, n, t = c
                base = t.replace("const ", "").rstrip("*").strip()
                if n in mut_names:
                    args.append(f"io_ptr<{base}>({n}, {n}_out, stream)")
                else:
                    args.append(f"in_ptr<const {base}>({n})")
            else:
                args.append(c[1])
        w(f"  return status({h['name']}({', '.join(args)}), \"{h['name']}\");")
        w("}")
        bind = ["ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()"]
        bind += [".Arg<ffi::AnyBuffer>()" for _ in ops] + [".Ret<ffi::AnyBuffer>()" for _ in muts]
        bind += [f".Attr<{a['type']}>(\"{a['name']}\")" for a in attrs]
        w(f"XLA_FFI_DEFINE_HANDLER_SYMBOL({h['symbol']}, {h['name']}_impl,")
        line = "    "
        for b in bind:
            if len(line) + len(b) > 128:
                w(line)
                line = "        "
            line += b
        w(line + ");")
        w("")
    w("#else")
    w("#error \"xla/ffi/api/ffi.h not found: add -I$(python -c 'import jax.ffi; print(jax.ffi.include_dir())')\"")
    w("#endif")
    return "\n".join(o) + "\n"


def table(handlers):
    return {"generated_by": "scripts/gen_ffi.py", "handlers": [
        {"name": h["name"], "symbol": h["symbol"],
         "operands": [{"name": x["name"], "mutable": x["mutable"], "ctype": x["ctype"]} for x in h["operands"]],
         "attrs": h["attrs"]} for h in handlers]}


def generate():
    structs, aliases, protos = parse_header(open(HEADER).read())
    handlers = plan(structs, aliases, protos)
    return emit_cc(handlers), json.dumps(table(handlers), indent=1) + "\n", handlers, protos


def main():
    cc, js, handlers, protos = generate()
    os.makedirs(os.path.dirname(OUT_CC), exist_ok=True)
    open(OUT_CC, "w").write(cc)
    open(OUT_JSON, "w").write(js)
    print(f"{len(handlers)} handlers of {len(protos)} prototypes -> {os.path.relpath(OUT_CC, ROOT)}, {os.path.relpath(OUT_JSON, ROOT)}")


if __name__ == "__main__":
    main()
