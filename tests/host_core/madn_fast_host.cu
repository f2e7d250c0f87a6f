// Host-side harness (TEST INFRASTRUCTURE): runs the __host__ __device__ rule functions of
// csrc/madn_fast.cuh and csrc/madn_core.cuh on the CPU over NumPy structure-of-arrays buffers so that
// tests/test_madn_fast_core.py can compare them with the oracle without a GPU.
#include <cstdint>
#include <cstring>
#include "../../exploring-muzero-on-dog_b200/csrc/madn_track.cuh"

using namespace dogstep;

static void load(const MadnGeom& g, int64_t i, const int8_t* board, const int8_t* cur, const int8_t* pins, const int8_t* reward,
                 const uint8_t* done, const int8_t* aset, MadnRegs& s, uint64_t occ_board[4]) {
  for (int p = 0; p < 4; ++p) { s.occ[p] = 0; s.pins[p] = 0xFFFFFFFFu; s.as[p] = 0; }
  for (int c = 0; c < g.total; ++c) {
    int v = board[i * g.total + c];
    if (v >= 0 && v < 4) s.occ[v] |= 1ull << c;
  }
  for (int p = 0; p < g.n; ++p) {
    memcpy(&s.pins[p], pins + (i * g.n + p) * 4, 4);
    uint64_t row = 0;
    memcpy(&row, aset + (i * g.n + p) * 6, 6);
    s.as[p] = row;
  }
  s.cur = cur[i]; s.done = done[i] != 0; s.reward = reward[i]; s.die = 0;
  for (int p = 0; p < 4; ++p) occ_board[p] = s.occ[p];
}

static void store(const MadnGeom& g, int64_t i, int8_t* board, int8_t* cur, int8_t* pins, int8_t* reward, uint8_t* done,
                  int8_t* aset, const MadnRegs& s) {
  for (int c = 0; c < g.total; ++c) {
    int v = -1;
    for (int p = 0; p < 4; ++p) if ((s.occ[p] >> c) & 1ull) v = p;
    board[i * g.total + c] = (int8_t)v;
  }
  for (int p = 0; p < g.n; ++p) {
    memcpy(pins + (i * g.n + p) * 4, &s.pins[p], 4);
    memcpy(aset + (i * g.n + p) * 6, &s.as[p], 6);
  }
  cur[i] = (int8_t)s.cur; done[i] = (uint8_t)s.done; reward[i] = (int8_t)s.reward;
}

extern "C" {

// mask_out[i] = 24-bit valid mask from the fast core, canon_out[i] = is_canonical4
int hostcore_det_valid_mask4(int64_t n, uint32_t rules, int compile_time_train_rules, const int8_t* board, const int8_t* cur,
                             const int8_t* pins, const int8_t* reward, const uint8_t* done, const int8_t* aset,
                             uint32_t* mask_out, uint8_t* canon_out, uint32_t* generic_mask_out) {
  dogstep_madn_cfg cfg{4, 0xF, 10, rules};
  MadnGeom g;
  if (madn_make_geom(&cfg, &g)) return -1;
  for (int64_t i = 0; i < n; ++i) {
    MadnRegs s; uint64_t ob[4];
    load(g, i, board, cur, pins, reward, done, aset, s, ob);
    canon_out[i] = is_canonical4(s, ob);
    int cp = 0;
    generic_mask_out[i] = madn_det_valid_mask(g, s);
    mask_out[i] = canon_out[i] ? det_valid_mask4(RuleSet<kRulesRuntime>{g.rules}, g, s, cp) : generic_mask_out[i];
  }
  return 0;
}

// in-place fast step with a valid action index (0..23) for canonical live games; others untouched (stepped[i] = 0)
int hostcore_det_step4(int64_t n, uint32_t rules, int8_t* board, int8_t* cur, int8_t* pins, int8_t* reward, uint8_t* done,
                       int8_t* aset, const int32_t* action, uint8_t* stepped, uint8_t* still_canon) {
  dogstep_madn_cfg cfg{4, 0xF, 10, rules};
  MadnGeom g;
  if (madn_make_geom(&cfg, &g)) return -1;
  for (int64_t i = 0; i < n; ++i) {
    MadnRegs s; uint64_t ob[4];
    load(g, i, board, cur, pins, reward, done, aset, s, ob);
    stepped[i] = 0; still_canon[i] = 0;
    if (!is_canonical4(s, ob) || s.done) continue;
    int cp = 0;
    RuleSet<kRulesRuntime> R{g.rules};
    uint32_t m = det_valid_mask4(R, g, s, cp);
    if (m == 0u) { det_no_step4(s); still_canon[i] = 1; }
    else {
      if (!((m >> action[i]) & 1u)) return -2;
      det_step4(R, s, cp, action[i]);
      still_canon[i] = 1;
    }
    stepped[i] = 1;
    store(g, i, board, cur, pins, reward, done, aset, s);
  }
  return 0;
}

// the track-coordinate rules of csrc/madn_track.cuh (training rule set only): covered[i] = track_from_regs accepted the state;
// mask_out[i] = its 24-bit valid mask
int hostcore_track_valid_mask(int64_t n, const int8_t* board, const int8_t* cur, const int8_t* pins, const int8_t* reward,
                              const uint8_t* done, const int8_t* aset, uint32_t* mask_out, uint8_t* covered) {
  dogstep_madn_cfg cfg{4, 0xF, 10, kTrainRules};
  MadnGeom g;
  if (madn_make_geom(&cfg, &g)) return -1;
  for (int64_t i = 0; i < n; ++i) {
    MadnRegs s; uint64_t ob[4];
    load(g, i, board, cur, pins, reward, done, aset, s, ob);
    Track4 t;
    covered[i] = track_from_regs(s, t);
    int cp = 0;
    mask_out[i] = covered[i] ? track_valid_mask(t, cp) : 0u;
  }
  return 0;
}

// in-place: regs -> track -> step (valid action) / no_step -> regs, for covered live games; others untouched (stepped[i] = 0)
int hostcore_track_step(int64_t n, int8_t* board, int8_t* cur, int8_t* pins, int8_t* reward, uint8_t* done, int8_t* aset,
                        const int32_t* action, uint8_t* stepped) {
  dogstep_madn_cfg cfg{4, 0xF, 10, kTrainRules};
  MadnGeom g;
  if (madn_make_geom(&cfg, &g)) return -1;
  for (int64_t i = 0; i < n; ++i) {
    MadnRegs s; uint64_t ob[4];
    load(g, i, board, cur, pins, reward, done, aset, s, ob);
    stepped[i] = 0;
    Track4 t;
    if (!track_from_regs(s, t) || s.done) continue;
    int cp = 0;
    const uint32_t m = track_valid_mask(t, cp);
    if (m == 0u) track_no_step(t);
    else {
      if (!((m >> action[i]) & 1u)) return -2;
      track_step(t, cp, action[i]);
    }
    MadnRegs r;
    track_to_regs(g, t, r);
    stepped[i] = 1;
    store(g, i, board, cur, pins, reward, done, aset, r);
  }
  return 0;
}

}  // extern "C"
