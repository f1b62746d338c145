// host_ops.cc — host-side entry points of the C ABI that need no GPU: the request blob
// packer (Addons>ConcatInputs) and the run-time size-expression evaluator
// (Addons>ShapeConstruct).
//
// Reference counterparts (reference root relative):
//   recom_concat_inputs        <-> ConcatInputsOp::Compute
//                                  (tensorflow_addons/custom_ops/concat_inputs/concat_inputs_ops.cc:42-77):
//                                  memcpy every host-produced input into ONE blob so that a
//                                  single H2D copy feeds the fused op.
//   recom_shape_construct_*    <-> ShapeConstructOp
//                                  (tensorflow_addons/custom_ops/shape_construct/shape_construct_ops.cc:34-153):
//                                  the reference prints each size expression as C, compiles it
//                                  with g++ at kernel construction (or substitutes with SymEngine
//                                  at run time); here the same expression strings are parsed once
//                                  and evaluated directly, with C `int` semantics.
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/recom_b200.h"

namespace recom {
recom_status set_error(recom_status st, const char* fmt, ...);  // capi.cu
}
using recom::set_error;

// ---------------------------------------------------------------------------------------
// size expressions
// ---------------------------------------------------------------------------------------
namespace {

enum Op : int32_t { PUSH_CONST, PUSH_SYM, ADD, SUB, MUL, DIV, MOD, NEG, MAX, MIN, POW, ABS };

struct Instr {
  int32_t op;
  int64_t arg;
};

struct Parser {
  const char* s;
  const std::vector<std::string>& symbols;
  std::vector<Instr>& code;
  std::string err;

  void skip() {
    while (*s && isspace((unsigned char)*s)) ++s;
  }
  bool fail(const std::string& m) {
    if (err.empty()) err = m;
    return false;
  }
  bool expr() {  // + -
    if (!term()) return false;
    for (;;) {
      skip();
      if (*s == '+' || *s == '-') {
        const char o = *s++;
        if (!term()) return false;
        code.push_back({o == '+' ? ADD : SUB, 0});
      } else {
        return true;
      }
    }
  }
  bool term() {  // * / %
    if (!unary()) return false;
    for (;;) {
      skip();
      if (*s == '*' && s[1] == '*') {  // SymEngine prints powers as x**2
        s += 2;
        if (!unary()) return false;
        code.push_back({POW, 0});
      } else if (*s == '*' || *s == '/' || *s == '%') {
        const char o = *s++;
        if (!unary()) return false;
        code.push_back({o == '*' ? MUL : (o == '/' ? DIV : MOD), 0});
      } else {
        return true;
      }
    }
  }
  bool unary() {
    skip();
    if (*s == '-') {
      ++s;
      if (!unary()) return false;
      code.push_back({NEG, 0});
      return true;
    }
    if (*s == '+') {
      ++s;
      return unary();
    }
    return atom();
  }
  bool atom() {
    skip();
    if (*s == '(') {
      ++s;
      if (!expr()) return false;
      skip();
      if (*s != ')') return fail("expected ')'");
      ++s;
      return true;
    }
    if (isdigit((unsigned char)*s)) {
      int64_t v = 0;
      while (isdigit((unsigned char)*s)) v = v * 10 + (*s++ - '0');
      code.push_back({PUSH_CONST, v});
      return true;
    }
    if (isalpha((unsigned char)*s) || *s == '_') {
      std::string name;
      while (isalnum((unsigned char)*s) || *s == '_') name += *s++;
      skip();
      if (*s == '(') {  // function call
        ++s;
        int nargs = 0;
        skip();
        if (*s != ')') {
          for (;;) {
            if (!expr()) return false;
            ++nargs;
            skip();
            if (*s == ',') {
              ++s;
              continue;
            }
            break;
          }
        }
        if (*s != ')') return fail("expected ')' after arguments of " + name);
        ++s;
        std::string f = name;
        std::transform(f.begin(), f.end(), f.begin(), [](unsigned char c) { return (char)tolower(c); });
        if ((f == "max" || f == "fmax") && nargs >= 2) {
          for (int i = 1; i < nargs; ++i) code.push_back({MAX, 0});
        } else if ((f == "min" || f == "fmin") && nargs >= 2) {
          for (int i = 1; i < nargs; ++i) code.push_back({MIN, 0});
        } else if (f == "pow" && nargs == 2) {
          code.push_back({POW, 0});
        } else if ((f == "abs" || f == "fabs") && nargs == 1) {
          code.push_back({ABS, 0});
        } else if ((f == "floor" || f == "ceiling" || f == "ceil") && nargs == 1) {
          // integer arithmetic: the argument is already an integer
        } else {
          return fail("unsupported function '" + name + "'");
        }
        return true;
      }
      for (size_t i = 0; i < symbols.size(); ++i)
        if (symbols[i] == name) {
          code.push_back({PUSH_SYM, (int64_t)i});
          return true;
        }
      return fail("unknown symbol '" + name + "'");
    }
    return fail(std::string("unexpected character '") + *s + "'");
  }
};

bool run(const std::vector<Instr>& code, const int64_t* in, int64_t* out) {
  int64_t st[64];
  int sp = 0;
  for (const Instr& i : code) {
    if (i.op == PUSH_CONST || i.op == PUSH_SYM) {
      if (sp >= 64) return false;
      // the reference binds every symbol as `const int` (shape_construct_ops.cc:71-75)
      st[sp++] = i.op == PUSH_CONST ? i.arg : (int64_t)(int32_t)in[i.arg];
      continue;
    }
    if (i.op == NEG || i.op == ABS) {
      if (sp < 1) return false;
      st[sp - 1] = i.op == NEG ? -st[sp - 1] : (st[sp - 1] < 0 ? -st[sp - 1] : st[sp - 1]);
      continue;
    }
    if (sp < 2) return false;
    const int64_t b = st[--sp], a = st[sp - 1];
    int64_t r = 0;
    switch (i.op) {
      case ADD: r = a + b; break;
      case SUB: r = a - b; break;
      case MUL: r = a * b; break;
      case DIV: if (b == 0) return false; r = a / b; break;  // C: truncation toward zero
      case MOD: if (b == 0) return false; r = a % b; break;
      case MAX: r = a > b ? a : b; break;
      case MIN: r = a < b ? a : b; break;
      case POW: {
        if (b < 0) return false;
        r = 1;
        for (int64_t k = 0; k < b; ++k) r *= a;
        break;
      }
      default: return false;
    }
    st[sp - 1] = r;
  }
  if (sp != 1) return false;
  *out = st[0];
  return true;
}

}  // namespace

struct recom_shape_construct {
  std::vector<std::string> symbols;
  std::vector<std::vector<Instr>> programs;
};

extern "C" {

recom_status recom_shape_construct_create(const char* const* symbols, int32_t num_symbols,
                                          const char* const* exprs, int32_t num_exprs,
                                          recom_shape_construct** out) {
  if (!out || num_symbols < 0 || num_exprs <= 0 || !exprs || (num_symbols > 0 && !symbols))
    return set_error(RECOM_ERR_INVALID_ARGUMENT, "null or empty argument");
  auto* sc = new recom_shape_construct();
  for (int32_t i = 0; i < num_symbols; ++i) sc->symbols.emplace_back(symbols[i] ? symbols[i] : "");
  for (int32_t e = 0; e < num_exprs; ++e) {
    std::vector<Instr> code;
    Parser p{exprs[e] ? exprs[e] : "", sc->symbols, code, {}};
    bool ok = p.expr();
    p.skip();
    if (ok && *p.s) ok = p.fail(std::string("trailing input at '") + p.s + "'");
    if (!ok) {
      const std::string msg = p.err;
      delete sc;
      return set_error(RECOM_ERR_INVALID_ARGUMENT, "expression %d ('%s'): %s", e, exprs[e] ? exprs[e] : "",
                       msg.c_str());
    }
    sc->programs.push_back(std::move(code));
  }
  *out = sc;
  return RECOM_OK;
}

recom_status recom_shape_construct_eval(const recom_shape_construct* sc, const int64_t* inputs,
                                        int32_t num_inputs, int64_t* outputs, int32_t num_outputs) {
  if (!sc || !outputs || (num_inputs > 0 && !inputs))
    return set_error(RECOM_ERR_INVALID_ARGUMENT, "null argument");
  if (num_inputs != (int32_t)sc->symbols.size() || num_outputs != (int32_t)sc->programs.size())
    return set_error(RECOM_ERR_INVALID_ARGUMENT, "expected %zu inputs and %zu outputs, got %d and %d",
                     sc->symbols.size(), sc->programs.size(), num_inputs, num_outputs);
  for (size_t e = 0; e < sc->programs.size(); ++e)
    if (!run(sc->programs[e], inputs, &outputs[e]))
      return set_error(RECOM_ERR_INVALID_ARGUMENT, "expression %zu: evaluation failed (division by zero?)", e);
  return RECOM_OK;
}

recom_status recom_shape_construct_destroy(recom_shape_construct* sc) {
  delete sc;
  return RECOM_OK;
}

// ---------------------------------------------------------------------------------------
// ConcatInputs
// ---------------------------------------------------------------------------------------
namespace {
// Shared body of the two packers. `elems` == nullptr: no argument trailer.
recom_status concat_impl(const void* const* inputs, const int64_t* nbytes, const int64_t* elems,
                         int32_t num_inputs, void* blob, size_t blob_bytes, int64_t* offsets,
                         int64_t* args_offset, size_t* needed) {
  if (num_inputs < 0 || !needed || (num_inputs > 0 && (!nbytes || !offsets)))
    return set_error(RECOM_ERR_INVALID_ARGUMENT, "null argument");
  size_t off = 0;
  for (int32_t i = 0; i < num_inputs; ++i) {
    if (nbytes[i] < 0) return set_error(RECOM_ERR_INVALID_ARGUMENT, "input %d: negative size", i);
    offsets[i] = (int64_t)off;
    off = (off + (size_t)nbytes[i] + 15) / 16 * 16;  // slots are 16-byte aligned and padded
  }
  const size_t data_end = std::max<size_t>((off + 15) / 16 * 16, 16);
  // trailer: {magic, num_slots, offsets[num_slots], elems[num_slots]}, 16-byte aligned
  const size_t trailer = elems ? (16 + 2 * (size_t)num_inputs * 8 + 15) / 16 * 16 : 0;
  const size_t total = data_end + trailer;
  *needed = total;
  if (args_offset) *args_offset = elems ? (int64_t)data_end : -1;
  if (!blob) return RECOM_OK;  // size query
  if (blob_bytes < total)
    return set_error(RECOM_ERR_WORKSPACE_TOO_SMALL, "blob needs %zu bytes, got %zu", total, blob_bytes);
  if (num_inputs > 0 && !inputs) return set_error(RECOM_ERR_INVALID_ARGUMENT, "null inputs");
  auto copy_range = [&](int32_t lo, int32_t hi) {
    for (int32_t i = lo; i < hi; ++i) {
      uint8_t* dst = (uint8_t*)blob + offsets[i];
      if (nbytes[i]) memcpy(dst, inputs[i], (size_t)nbytes[i]);
      const size_t end = (size_t)offsets[i] + (size_t)nbytes[i];
      const size_t pad_to = i + 1 < num_inputs ? (size_t)offsets[i + 1] : data_end;
      if (pad_to > end) memset((uint8_t*)blob + end, 0, pad_to - end);
    }
  };
  // the reference copies on one thread (concat_inputs_ops.cc:62-70); large requests are split
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const int32_t nthreads = (int32_t)std::min<size_t>(std::min<unsigned>(hw, 8), total / (1 << 20) + 1);
  if (nthreads <= 1 || num_inputs < 2 * nthreads) {
    copy_range(0, num_inputs);
  } else {
    std::vector<std::thread> th;
    for (int32_t t = 0; t < nthreads; ++t)
      th.emplace_back(copy_range, (int32_t)((int64_t)num_inputs * t / nthreads),
                      (int32_t)((int64_t)num_inputs * (t + 1) / nthreads));
    for (auto& x : th) x.join();
  }
  if (elems) {
    uint8_t* t = (uint8_t*)blob + data_end;
    memset(t, 0, trailer);
    const uint64_t magic = 0x5247524132304252ull;  // kArgsMagic (csrc/plan.h)
    const int64_t n = num_inputs;
    memcpy(t, &magic, 8);
    memcpy(t + 8, &n, 8);
    if (num_inputs) {
      memcpy(t + 16, offsets, (size_t)num_inputs * 8);
      memcpy(t + 16 + (size_t)num_inputs * 8, elems, (size_t)num_inputs * 8);
    }
  }
  return RECOM_OK;
}
}  // namespace

recom_status recom_concat_inputs(const void* const* inputs, const int64_t* nbytes, int32_t num_inputs,
                                 void* blob, size_t blob_bytes, int64_t* offsets, size_t* needed) {
  return concat_impl(inputs, nbytes, nullptr, num_inputs, blob, blob_bytes, offsets, nullptr, needed);
}

recom_status recom_concat_inputs_with_args(const void* const* inputs, const int64_t* nbytes,
                                           const int64_t* elems, int32_t num_inputs, void* blob,
                                           size_t blob_bytes, int64_t* offsets, int64_t* args_offset,
                                           size_t* needed) {
  if (num_inputs > 0 && !elems) return set_error(RECOM_ERR_INVALID_ARGUMENT, "null elems");
  static const int64_t kNone = 0;
  return concat_impl(inputs, nbytes, elems ? elems : &kNone, num_inputs, blob, blob_bytes, offsets,
                     args_offset, needed);
}

}  // extern "C"
