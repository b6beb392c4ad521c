// ctd_layout_hot.h -- ptxas places the device functions of a kernel in the order of their mangled names.  Renaming the functions
// the playout loop runs every step to equally long names with a common prefix puts them next to each other (one contiguous
// region that maps evenly onto the instruction cache's sets) instead of interleaved with the once-per-game code.
#pragma once
#define ctd_philox ctd_h00_philox____________________
#define ctd_has ctd_h01_has_______________________
#define ctd_append ctd_h02_append____________________
#define ctd_draw ctd_h03_draw______________________
#define ctd_take_like ctd_h04_take_like_________________
#define ctd_count_type ctd_h05_count_type________________
#define ctd_count_suit ctd_h06_count_suit________________
#define ctd_player_from_rank ctd_h07_player_from_rank__________
#define ctd_setup_next_player ctd_h08_setup_next_player_________
#define ctd_refresh_used_roles ctd_h09_refresh_used_roles________
#define ctd_apply_finish ctd_h10_apply_finish______________
#define ctd_apply_build ctd_h11_apply_build_______________
#define ctd_move_crown ctd_h12_move_crown________________
#define ctd_check_game_ending ctd_h13_check_game_ending_________
#define ctd_setup_round ctd_h14_setup_round_______________
#define ctd_shuffle_bytes ctd_h15_shuffle_bytes_____________
#define ctd_reshuffle_if_empty ctd_h16_reshuffle_if_empty________
#define ctd_apply ctd_h17_apply_____________________
#define ctd_warp_choose ctd_h18_warp_choose_______________
