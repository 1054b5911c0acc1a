"""Host-side mirror of the reference's sampler selection (``inference/sampler/Sampler.java:74-122``).

The reference copies the logits row to the host whenever ``temperature > 0`` and samples there
(``CategoricalSampler.java:28-40``, ``ToppSampler.java:62-156``).  Here the arithmetic runs on the device
(``csrc/sampler.cuh`` behind ``b200_forward_decode_sample``): per token 4 bytes go in -- the uniform number -- and 4 bytes
come out.  What stays on the host is the random stream itself, because it is host state in the reference too:
``RandomGeneratorFactory.getDefault().create(seed)`` = ``L32X64MixRandom`` (Sampler.java:84), restated below from the published
LXM algorithm (Steele & Vigna, OOPSLA 2021).  No JDK is available here, so the stream is unpinned against a JVM; the test suite
checks this restatement against the oracle's independent C restatement.
"""
from __future__ import annotations

M32 = 0xFFFFFFFF


def _mix_murmur32(z: int) -> int:
    z = ((z ^ (z >> 16)) * 0x85EBCA6B) & M32
    z = ((z ^ (z >> 13)) * 0xC2B2AE35) & M32
    return z ^ (z >> 16)


def _mix_lea32(z: int) -> int:
    z = ((z ^ (z >> 16)) * 0xD36D884B) & M32
    z = ((z ^ (z >> 16)) * 0xD36D884B) & M32
    return z ^ (z >> 16)


def _rotl32(v: int, k: int) -> int:
    return ((v << k) | (v >> (32 - k))) & M32


class L32X64MixRandom:
    """java.util.random default generator: 32-bit LCG (multiplier 0xadb4a92d) + xoroshiro64, mixed with mixLea32."""

    def __init__(self, seed: int):
        sd = (seed ^ 0x6A09E667F3BCC909) & 0xFFFFFFFFFFFFFFFF  # SILVER_RATIO_64
        self.a = _mix_murmur32(sd >> 32) | 1
        self.s = 1
        self.x0 = _mix_lea32(sd & M32)
        self.x1 = _mix_lea32(((sd & M32) + 0x9E3779B9) & M32)  # GOLDEN_RATIO_32
        if (self.x0 | self.x1) == 0:
            self.x0, self.x1 = 0x9E3779B9, 0x3C6EF372

    def next_int(self) -> int:
        result = _mix_lea32((self.s + self.x0) & M32)
        self.s = (0xADB4A92D * self.s + self.a) & M32
        q0, q1 = self.x0, self.x1
        q1 ^= q0
        q0 = _rotl32(q0, 26)
        q0 = q0 ^ q1 ^ ((q1 << 9) & M32)
        q1 = _rotl32(q1, 13)
        self.x0, self.x1 = q0, q1
        return result

    def next_float1(self) -> float:
        """RandomGenerator.nextFloat(1f): (nextInt() >>> 8) * 2^-24 (times the bound 1, always below it)."""
        return (self.next_int() >> 8) * (2.0 ** -24)


class Sampler:
    """``Sampler.selectSampler(vocabularySize, temperature, topp, rngSeed)``: greedy when temperature == 0, otherwise
    temperature + softmax + categorical (topp outside (0,1)) or top-p sampling.  ``sample_token`` runs one decode forward and
    returns the sampled id; the logits never leave the GPU."""

    def __init__(self, vocabulary_size: int, temperature: float, topp: float, rng_seed: int):
        self.vocabulary_size = vocabulary_size
        self.temperature = float(temperature)
        self.topp = float(topp)
        self.rng = None if self.temperature == 0.0 else L32X64MixRandom(rng_seed)

    def sample_token(self, plan, token: int, position: int) -> int:
        if self.temperature == 0.0:
            return plan.forward_decode(token, position, logits=False)[1]
        return plan.forward_decode_sample(token, position, self.temperature, self.topp, self.rng.next_float1())


def select_sampler(vocabulary_size: int, temperature: float, topp: float, rng_seed: int) -> Sampler:
    return Sampler(vocabulary_size, temperature, topp, rng_seed)
