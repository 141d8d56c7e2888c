//! `src/bin/parity_dump.rs` of the patched reference crate: runs ONE agent of one configuration on the injected stream
//! and prints what `Agent::train` returns, one JSON object on stdout, for tests/test_rust_ref.py to hold the oracle to.
//!
//! usage: parity_dump ENV AGENT SELECTOR POLICY TARGET SEED AGENT_ID N_EPISODES EVAL_AT [MAX_STEPS SLIPPERY MAP]
//!   ENV 0 blackjack | 1 frozen_lake | 2 cliff_walking | 3 taxi;  AGENT 0 one-step | 1 traces;  SELECTOR 0 eps-greedy | 1 ucb
//!   POLICY 0 basic | 1 double;  TARGET 0 sarsa | 1 qlearning | 2 expected_sarsa;  MAP 4x4 | 8x8 (frozen_lake)
//! Hyper-parameters are the bins' defaults (bin/taxi.rs:22-68) with epsilon_decay = 1 / (0.5 * N_EPISODES) (:78).
//! f64 values are printed as their bit patterns (u64), so the comparison is bit for bit.
use std::rc::Rc;

use reinforcement_learning::action_selection::{EnumActionSelection, UniformEpsilonGreed, UpperConfidenceBound};
use reinforcement_learning::agent::{expected_sarsa, qlearning, sarsa, Agent, ElegibilityTracesAgent, GetNextQValue, OneStepAgent};
use reinforcement_learning::env::{BlackJackEnv, CliffWalkingEnv, Env, FrozenLakeEnv, TaxiEnv};
use reinforcement_learning::policy::{DoubleTabularPolicy, EnumPolicy, TabularPolicy};
use reinforcement_learning::rng;

struct Cfg {
    agent: u32,
    selector: u32,
    policy: u32,
    target: u32,
    n_episodes: u128,
    eval_at: u128,
}

fn bits(v: &[f64]) -> String {
    let parts: Vec<String> = v.iter().map(|x| x.to_bits().to_string()).collect();
    format!("[{}]", parts.join(","))
}

fn run<const COUNT: usize>(env: &mut dyn Env<usize, COUNT>, cfg: &Cfg) {
    let learning_rate: f64 = 0.05;
    let discount_factor: f64 = 0.95;
    let lambda_factor: f64 = 0.5;
    let epsilon_decay: f64 = 1.0 / (0.5 * cfg.n_episodes as f64);
    let selector: EnumActionSelection<usize, COUNT> = if cfg.selector == 0 {
        EnumActionSelection::from(UniformEpsilonGreed::new(1.0, Rc::new(move |a| a - epsilon_decay), 0.0))
    } else {
        EnumActionSelection::from(UpperConfidenceBound::new(0.5))
    };
    let policy: EnumPolicy<usize, COUNT> = if cfg.policy == 0 {
        EnumPolicy::from(TabularPolicy::new(learning_rate, 0.0))
    } else {
        EnumPolicy::from(DoubleTabularPolicy::new(learning_rate, 0.0))
    };
    let func: GetNextQValue<COUNT> = match cfg.target {
        0 => sarsa,
        1 => qlearning,
        _ => expected_sarsa,
    };
    let (reward_history, episode_length, training_error) = if cfg.agent == 0 {
        let mut agent: OneStepAgent<usize, COUNT> = OneStepAgent::new(policy, discount_factor, selector, func);
        agent.train(env, cfg.n_episodes, cfg.eval_at)
    } else {
        let mut agent: ElegibilityTracesAgent<usize, COUNT> =
            ElegibilityTracesAgent::new(policy, discount_factor, selector, lambda_factor, func);
        agent.train(env, cfg.n_episodes, cfg.eval_at)
    };
    let lens: Vec<String> = episode_length.iter().map(|x| x.to_string()).collect();
    println!(
        "{{\"reward_history_bits\":{},\"episode_length\":[{}],\"training_error_bits\":{},\"rng_words\":{}}}",
        bits(&reward_history),
        lens.join(","),
        bits(&training_error),
        rng::position()
    );
}

fn main() {
    let a: Vec<String> = std::env::args().collect();
    if a.len() < 10 {
        eprintln!("usage: parity_dump ENV AGENT SELECTOR POLICY TARGET SEED AGENT_ID N_EPISODES EVAL_AT [MAX_STEPS SLIPPERY MAP]");
        std::process::exit(2);
    }
    let p = |i: usize| -> u64 {
        let s = &a[i];
        if let Some(h) = s.strip_prefix("0x") { u64::from_str_radix(h, 16).unwrap() } else { s.parse::<u64>().unwrap() }
    };
    let cfg = Cfg { agent: p(2) as u32, selector: p(3) as u32, policy: p(4) as u32, target: p(5) as u32, n_episodes: p(8) as u128, eval_at: p(9) as u128 };
    let max_steps: u128 = if a.len() > 10 { p(10) as u128 } else { 100 };
    let slippery: bool = a.len() > 11 && p(11) != 0;
    let map_name: &str = if a.len() > 12 { a[12].as_str() } else { "8x8" };
    // the stream starts BEFORE the env is built: BlackJackEnv::new() deals a hand (env/blackjack.rs:57)
    rng::seed(p(6), p(7));
    match p(1) {
        0 => run::<2>(&mut BlackJackEnv::new(), &cfg),
        1 => {
            let map: &[&str] = if map_name == "4x4" { &FrozenLakeEnv::MAP_4X4 } else { &FrozenLakeEnv::MAP_8X8 };
            run::<4>(&mut FrozenLakeEnv::new(map, slippery, max_steps), &cfg)
        }
        2 => run::<4>(&mut CliffWalkingEnv::new(max_steps), &cfg),
        _ => run::<6>(&mut TaxiEnv::new(max_steps), &cfg),
    }
}
