// Test-only host shim: compiles the product's __host__ __device__ bitboard header with g++
// so its logic can be checked on the CPU against the oracle / golden vectors (no GPU needed).
#include "../othellozero_b200/csrc/oz_bitboard.cuh"
extern "C" {
unsigned long long bbh_legal(unsigned long long own, unsigned long long opp, int n) {
    return ozbb::legal_moves(own, opp, ozbb::full_mask(n));
}
unsigned long long bbh_flip(int sq, unsigned long long own, unsigned long long opp) {
    return ozbb::flip_mask(1ull << sq, own, opp);
}
unsigned long long bbh_legal_compact(unsigned long long own, unsigned long long opp, int n) {
    return ozbb::legal_moves_compact(own, opp, ozbb::full_mask(n));
}
unsigned long long bbh_flip_compact(int sq, unsigned long long own, unsigned long long opp) {
    return ozbb::flip_mask_compact(1ull << sq, own, opp);
}
unsigned bbh_play(int sq, unsigned long long* own, unsigned long long* opp, int n, unsigned long long* next_legal) {
    return ozbb::play_move(1ull << sq, own, opp, ozbb::full_mask(n), next_legal);
}
int bbh_kth(unsigned long long x, int k) { return ozbb::kth_set_bit(x, k); }
unsigned long long bbh_sm64(unsigned long long x) { return ozbb::sm64(x); }
void bbh_initial(int n, unsigned long long* b, unsigned long long* w) { ozbb::initial_position(n, b, w); }
}
