/* dogstep_rules.h — rule-set encoding shared by the C-ABI, the host mirror and the
 * test oracle.  Pure constants, no code.
 *
 * The reference keeps its rule switches as a static Python dict on every env pytree
 * (MADN/deterministic_madn.py:109-119, MADN/classic_madn.py:119-130,
 * DOG/dog.py rules dict built in env_reset :83-181).  XLA bakes one program per
 * dict; here the dict becomes one uint32 bitmask handed to every entry point.
 */
#ifndef DOGSTEP_RULES_H
#define DOGSTEP_RULES_H

#define DOGSTEP_RULE_TEAMS               (1u << 0)  /* enable_teams (forced off unless num_players==4) */
#define DOGSTEP_RULE_INITIAL_FREE_PIN    (1u << 1)  /* enable_initial_free_pin */
#define DOGSTEP_RULE_CIRCULAR_BOARD      (1u << 2)  /* enable_circular_board */
#define DOGSTEP_RULE_START_BLOCKING      (1u << 3)  /* enable_start_blocking */
#define DOGSTEP_RULE_JUMP_IN_GOAL        (1u << 4)  /* enable_jump_in_goal_area */
#define DOGSTEP_RULE_FRIENDLY_FIRE       (1u << 5)  /* enable_friendly_fire */
#define DOGSTEP_RULE_START_ON_1          (1u << 6)  /* enable_start_on_1 (MADN only) */
#define DOGSTEP_RULE_BONUS_TURN_ON_6     (1u << 7)  /* enable_bonus_turn_on_6 (MADN only) */
#define DOGSTEP_RULE_MUST_TRAVERSE_START (1u << 8)  /* must_traverse_start */
#define DOGSTEP_RULE_DICE_RETHROW        (1u << 9)  /* enable_dice_rethrow (classic MADN only) */
#define DOGSTEP_RULE_DISABLE_SWAPPING    (1u << 10) /* DOG: disable_swapping */
#define DOGSTEP_RULE_DISABLE_HOT_SEVEN   (1u << 11) /* DOG: disable_hot_seven */
#define DOGSTEP_RULE_DISABLE_JOKER       (1u << 12) /* DOG: disable_joker */

/* error codes returned by every dogstep_* entry point (0 == success) */
#define DOGSTEP_OK                0
#define DOGSTEP_ERR_INVALID_ARG  -1   /* null pointer, n < 0, bad num_players/distance/layout */
#define DOGSTEP_ERR_UNSUPPORTED  -2   /* configuration outside what the kernels implement */
#define DOGSTEP_ERR_CUDA         -3   /* launch/runtime failure; see dogstep_last_error() */

#endif /* DOGSTEP_RULES_H */
