// Host-side harness (TEST INFRASTRUCTURE): runs the __host__ __device__ DOG rule functions of csrc/dog_fast.cuh on
// the CPU over NumPy structure-of-arrays buffers so tests/test_dog_fast_core.py can compare them with the oracle.
#include <cstdint>
#include <cstring>
#include "../../exploring-muzero-on-dog_b200/csrc/dog_fast.cuh"

using namespace dogstep;

struct Soa {
  int8_t* board; int8_t* cur; int32_t* pins; int8_t* reward; uint8_t* done; int8_t* deck; int8_t* hands;
  int8_t* swap_choices; int8_t* round_starter; int8_t* phase; uint32_t* key; int8_t* hand_size;
};

static void load(const Soa& a, int64_t i, DogS& s) {
  memset(&s, 0, sizeof(s));
  for (int k = 0; k < 64; ++k) s.board[k] = k < 56 ? a.board[i * 56 + k] : (int8_t)-1;
  for (int p = 0; p < 4; ++p)
    for (int k = 0; k < 4; ++k) s.pins[p][k] = a.pins[(i * 4 + p) * 4 + k];
  for (int p = 0; p < 4; ++p)
    for (int k = 0; k < 14; ++k) s.hands[p][k] = a.hands[(i * 4 + p) * 14 + k];
  for (int k = 0; k < 14; ++k) s.deck[k] = a.deck[i * 14 + k];
  for (int k = 0; k < 4; ++k) s.swap_choices[k] = a.swap_choices[i * 4 + k];
  s.key[0] = a.key[2 * i]; s.key[1] = a.key[2 * i + 1];
  s.cur = a.cur[i]; s.reward = a.reward[i]; s.done = a.done[i] != 0; s.round_starter = a.round_starter[i];
  s.phase = a.phase[i]; s.hand_size = a.hand_size[i];
}

static void store(const Soa& a, int64_t i, const DogS& s) {
  for (int k = 0; k < 56; ++k) a.board[i * 56 + k] = s.board[k];
  for (int p = 0; p < 4; ++p)
    for (int k = 0; k < 4; ++k) a.pins[(i * 4 + p) * 4 + k] = s.pins[p][k];
  for (int p = 0; p < 4; ++p)
    for (int k = 0; k < 14; ++k) a.hands[(i * 4 + p) * 14 + k] = s.hands[p][k];
  a.cur[i] = (int8_t)s.cur; a.reward[i] = (int8_t)s.reward; a.done[i] = (uint8_t)s.done;
}

extern "C" {

// valid_actions of the play phase through the fast functions: mask[n, 806]; canon[n]
int hostcore_dog_mask4(int64_t n, uint32_t rules, const Soa* a, uint8_t* mask, uint8_t* canon) {
  const Dog4Rules R = dg4_rules(rules);
  for (int64_t i = 0; i < n; ++i) {
    DogS s;
    load(*a, i, s);
    uint8_t* m = mask + i * 806;
    memset(m, 0, 806);
    canon[i] = dg4_canonical(s.pins, s.board, s.cur);
    if (!canon[i]) continue;
    if (s.phase != 0) {
      for (int c = 0; c < 14; ++c) m[792 + c] = s.hands[s.cur][c] > 0;
      continue;
    }
    Dog4View v;
    dg4_view(R, s.pins, s.cur, v);
    uint32_t pin_ok; uint64_t cell_ok;
    dg4_val_swap(R, v, pin_ok, cell_ok);
    const int8_t* hand = s.hands[v.cp];
    for (int b = 0; b < 396; ++b) {
      const int card = dg4_card_of_base(b);
      if (!(hand[0] > 0 || hand[card] > 0)) continue;
      if (!dg4_base_valid(R, v, b, pin_ok, cell_ok)) continue;
      if (hand[0] > 0) m[b] = 1;
      if (hand[card] > 0) m[396 + b] = 1;
    }
  }
  return 0;
}

// play-phase env_step through the fast functions for canonical, play-phase games; board rebuilt from the pins.
// stepped[i] = 1 when handled; deal[i] = 1 when the reference would deal next (not done here)
int hostcore_dog_play_phase4(int64_t n, uint32_t rules, const Soa* a, const int32_t* action, uint8_t* stepped, uint8_t* deal) {
  const Dog4Rules R = dg4_rules(rules);
  dogstep_madn_cfg cfg{4, 0xF, 10, rules};
  DogGeom g;
  if (dog_make_geom(&cfg, &g)) return -1;
  for (int64_t i = 0; i < n; ++i) {
    DogS s;
    load(*a, i, s);
    stepped[i] = 0; deal[i] = 0;
    if (s.phase != 0 || !dg4_canonical(s.pins, s.board, s.cur) || action[i] < 0 || action[i] >= 792) continue;
    int r, d;
    deal[i] = (uint8_t)dg4_play_phase(R, s, action[i], r, d);
    dog_set_pins_on_board(g, s.pins, s.board);
    stepped[i] = 1;
    store(*a, i, s);
  }
  return 0;
}

}  // extern "C"
