"""IncDiv of the weight-gradient kernel (omni-pq_b200/csrc/mlp_gemm_tc.cu) restated: the quotients position / nsample,
position / (npoint * nsample) and position / group are carried from k-block to k-block (positions advance by 32)
instead of being divided out per position; `q_at(16)` serves the warp's second position of a k-block."""
import numpy as np


class IncDiv:
    def __init__(self, r, d, step):
        self.d = d if d > 0 else 1
        self.q, self.rem = divmod(r, self.d)
        self.dq, self.dr = divmod(step, self.d)

    def advance(self):
        self.q += self.dq
        self.rem += self.dr
        if self.rem >= self.d:
            self.rem -= self.d
            self.q += 1

    def q_at(self, delta):
        q, r = self.q, self.rem + delta
        while r >= self.d:
            r -= self.d
            q += 1
        return q


def test_incremental_quotients_equal_division():
    rng = np.random.default_rng(0)
    for _ in range(400):
        d = int(rng.choice([1, 2, 3, 5, 8, 15, 16, 17, 31, 32, 33, 64, 100, 2048 * 64, 1 << 20]))
        r0 = int(rng.integers(0, 1 << 22))
        inc = IncDiv(r0, d, 32)
        r = r0
        for _ in range(int(rng.integers(1, 200))):
            assert inc.q == r // d and inc.rem == r % d
            assert inc.q_at(16) == (r + 16) // d
            inc.advance()
            r += 32
