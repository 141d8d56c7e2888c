//! Injected RNG stream — becomes `src/rng.rs` of the reference crate (see README.md in this directory).
//!
//! The crate draws every random number from `rand::thread_rng()`, which cannot be seeded.  For seed-for-seed parity
//! runs this module supplies a per-thread Philox4x32-10 word stream with the same role: `crate::rng::stream()` returns
//! a zero-sized `RngCore` handle, and the nine live call sites use it in place of `rand::thread_rng()`
//! (reference_rng.patch).  The stream is the contract of SURVEY.md §8.2 / include/rlb.h:
//!     w[n] = Philox4x32-10(key = seed, ctr = (n >> 2, agent_id))[n & 3],  next_u64 = w[n] | w[n+1] << 32
//! rand 0.8.5's `Uniform<f64|usize|u8>` and `gen_range` then map the words to values exactly as they do for ThreadRng.
use rand::{Error, RngCore};
use std::cell::RefCell;

#[derive(Clone, Debug)]
pub struct PhiloxStream {
    key: [u32; 2],
    agent: u64,
    n: u64,
}

impl PhiloxStream {
    pub fn new(seed: u64, agent: u64) -> Self {
        Self { key: [seed as u32, (seed >> 32) as u32], agent, n: 0 }
    }

    /// index of the next 32-bit word (rlb_agent_state.rng_n on the engine side)
    pub fn position(&self) -> u64 {
        self.n
    }

    fn block(&self, blk: u64) -> [u32; 4] {
        let mut c: [u32; 4] = [blk as u32, (blk >> 32) as u32, self.agent as u32, (self.agent >> 32) as u32];
        let mut k: [u32; 2] = self.key;
        for _ in 0..10 {
            let m0: u64 = 0xD251_1F53u64 * c[0] as u64;
            let m1: u64 = 0xCD9E_8D57u64 * c[2] as u64;
            c = [
                ((m1 >> 32) as u32) ^ c[1] ^ k[0],
                m1 as u32,
                ((m0 >> 32) as u32) ^ c[3] ^ k[1],
                m0 as u32,
            ];
            k[0] = k[0].wrapping_add(0x9E37_79B9);
            k[1] = k[1].wrapping_add(0xBB67_AE85);
        }
        c
    }
}

impl RngCore for PhiloxStream {
    fn next_u32(&mut self) -> u32 {
        let w: u32 = self.block(self.n >> 2)[(self.n & 3) as usize];
        self.n += 1;
        w
    }

    fn next_u64(&mut self) -> u64 {
        let lo: u64 = self.next_u32() as u64; // low word first, as rand_core's block RNGs do
        let hi: u64 = self.next_u32() as u64;
        lo | (hi << 32)
    }

    fn fill_bytes(&mut self, dest: &mut [u8]) {
        for chunk in dest.chunks_mut(4) {
            let w: [u8; 4] = self.next_u32().to_le_bytes();
            let len: usize = chunk.len();
            chunk.copy_from_slice(&w[..len]);
        }
    }

    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), Error> {
        self.fill_bytes(dest);
        Ok(())
    }
}

thread_local! {
    static STREAM: RefCell<PhiloxStream> = RefCell::new(PhiloxStream::new(0, 0));
}

/// (Re)start this thread's stream: `seed` is the Philox key, `agent` the global agent id in the counter's high words.
pub fn seed(seed: u64, agent: u64) {
    STREAM.with(|s| *s.borrow_mut() = PhiloxStream::new(seed, agent));
}

/// Words consumed so far by this thread's stream.
pub fn position() -> u64 {
    STREAM.with(|s| s.borrow().position())
}

/// The drop-in for `rand::thread_rng()`: a handle onto the thread-local stream.
pub fn stream() -> InjectedRng {
    InjectedRng
}

#[derive(Clone, Copy, Debug, Default)]
pub struct InjectedRng;

impl RngCore for InjectedRng {
    fn next_u32(&mut self) -> u32 {
        STREAM.with(|s| s.borrow_mut().next_u32())
    }

    fn next_u64(&mut self) -> u64 {
        STREAM.with(|s| s.borrow_mut().next_u64())
    }

    fn fill_bytes(&mut self, dest: &mut [u8]) {
        STREAM.with(|s| s.borrow_mut().fill_bytes(dest))
    }

    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), Error> {
        STREAM.with(|s| s.borrow_mut().try_fill_bytes(dest))
    }
}

#[cfg(test)]
mod tests {
    use super::*;

    // Random123 known-answer vectors (the same ones tests/test_oracle_kat.py pins the oracle to)
    #[test]
    fn philox_kat() {
        let s = PhiloxStream { key: [0, 0], agent: 0, n: 0 };
        assert_eq!(s.block(0), [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]);
        let s = PhiloxStream { key: [0xffffffff, 0xffffffff], agent: 0xffffffff_ffffffff, n: 0 };
        assert_eq!(s.block(0xffffffff_ffffffff), [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]);
        let s = PhiloxStream { key: [0xa4093822, 0x299f31d0], agent: 0x03707344_13198a2e, n: 0 };
        assert_eq!(s.block(0x85a308d3_243f6a88), [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]);
    }
}
