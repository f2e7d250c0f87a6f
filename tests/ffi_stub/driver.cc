// Compiles the GENERATED jax.ffi shim against the stand-in header and drives a few handlers through fake call frames
// (no GPU: only paths that return before a launch).  Prints one line per check; exit code 0 = all good.
#include <cstdio>
#include <cstring>
#include "../../exploring-muzero-on-dog_b200/csrc/ffi/dogstep_ffi.cc"

using xla::ffi::AnyBuffer;
using xla::ffi::CallFrame;

static int fails = 0;
static void expect(bool ok, const char* what) {
  std::printf("%s %s\n", ok ? "ok  " : "FAIL", what);
  fails += !ok;
}

static CallFrame frame_for(const xla::ffi::StubHandler& h) {
  CallFrame f;
  f.args.assign(h.num_args, AnyBuffer());   // empty buffers = NULL leaves
  f.rets.assign(h.num_rets, AnyBuffer());
  return f;
}

int main() {
  {  // valid_action of the deterministic game: negative n, bad player count, null leaves -> InvalidArgument, not a crash
    auto& h = k_dogstep_madn_det_valid_action;
    CallFrame f = frame_for(h);
    f.attrs = {{"n", int64_t(-1)}, {"cfg_num_players", int32_t(4)}, {"cfg_layout_mask", int32_t(15)}, {"cfg_distance", int32_t(10)},
               {"cfg_rules", uint32_t(0)}};
    auto e = h.call(f);
    expect(!e.success() && e.message().find("invalid argument") != std::string::npos, "madn_det_valid_action(n = -1) -> InvalidArgument");
    f.attrs["n"] = int64_t(4);
    f.attrs["cfg_distance"] = int32_t(13);
    e = h.call(f);
    expect(!e.success() && e.message().find("unsupported") != std::string::npos, "madn_det_valid_action(distance 13) -> unsupported");
    f.attrs["cfg_distance"] = int32_t(10);
    e = h.call(f);
    expect(!e.success(), "madn_det_valid_action(null leaves) -> InvalidArgument");
    expect(h.num_args == 8 && h.num_rets == 8 && h.attr_names.size() == 5, "madn_det_valid_action: 7 leaves + mask as operand and aliased result, 5 attributes");
  }
  {  // host keys travel as two uint32 attributes
    auto& h = k_dogstep_madn_det_play_random;
    bool has0 = false, has1 = false;
    for (auto& a : h.attr_names) { has0 |= a == "host_rng_key_0"; has1 |= a == "host_rng_key_1"; }
    expect(has0 && has1, "madn_det_play_random: host_rng_key as attributes host_rng_key_0 / _1");
  }
  {  // the search: the tree's 19 leaves as operands + aliased results
    auto& h = k_dogstep_mcts_expand_select;
    expect(h.num_rets >= 19 && h.num_args > h.num_rets, "mcts_expand_select: tree leaves aliased in place, network outputs read-only");
    CallFrame f = frame_for(h);
    for (auto& a : h.attr_names) f.attrs[a] = a == "n" ? xla::ffi::AttrValue(int64_t(8)) : (a.find("cfg_q_") == 0 || a.find("cfg_value") == 0 ||
        a.find("cfg_maxvisit") == 0 || a.find("cfg_eps") == 0 || a.find("cfg_pb") == 0 || a.find("cfg_dir") == 0 || a.find("cfg_temp") == 0 ||
        a.find("cfg_gumbel") == 0 ? xla::ffi::AttrValue(1.0f) : xla::ffi::AttrValue(int32_t(1)));
    auto e = h.call(f);
    expect(!e.success(), "mcts_expand_select(null tree) -> InvalidArgument");
  }
  {  // replay save: scalar fields of the array structs are attributes
    auto& h = k_dogstep_replay_save;
    bool cap = false;
    for (auto& a : h.attr_names) cap |= a == "buf_capacity";
    expect(cap, "replay_save: buf_capacity attribute");
  }
  std::printf("%d failures\n", fails);
  return fails ? 1 : 0;
}
