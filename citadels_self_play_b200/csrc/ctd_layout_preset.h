// ctd_layout_preset.h -- placement of the per-step device functions of ctd_k_playout_preset.
// ptxas lays the device functions of a kernel out in the order of their mangled names.  The playout kernels are bound by the SM's
// instruction cache (profiles/README.md), and which of their hot lines share cache sets is worth +-5 %: tools/layout_search.py
// timed 230 orders (random, then hill-climbing on the best) of the 19 functions the loop runs every step (renamed here to equally long names with an order prefix, so that
// they sit in one block in exactly this order, apart from the once-per-game code); this is the best one found for this unit
// (preset 1.076e9 with the functions where their own names put them -> 1.144e9 env steps/s).  Regenerate with the tool after changing the rules code.
#pragma once
#define ctd_has ctd_h00_ha
#define ctd_apply ctd_h01_ap
#define ctd_draw ctd_h02_dr
#define ctd_take_like ctd_h03_ta
#define ctd_setup_round ctd_h04_se
#define ctd_count_suit ctd_h05_co
#define ctd_refresh_used_roles ctd_h06_re
#define ctd_setup_next_player ctd_h07_se
#define ctd_check_game_ending ctd_h08_ch
#define ctd_apply_finish ctd_h09_ap
#define ctd_append ctd_h10_ap
#define ctd_count_type ctd_h11_co
#define ctd_philox ctd_h12_ph
#define ctd_player_from_rank ctd_h13_pl
#define ctd_warp_choose ctd_h14_wa
#define ctd_shuffle_bytes ctd_h15_sh
#define ctd_apply_build ctd_h16_ap
#define ctd_move_crown ctd_h17_mo
#define ctd_reshuffle_if_empty ctd_h18_re
