// include/rlb.hpp — header-only C++ mirror of the reference's trait surface for the hot path, over the C ABI
// of rlb.h.  The reference is a Rust crate and Rust is not available in the build image, so this is the
// compiled-language host side: same type names, constructor arguments, method names and error behaviour as
// JohnVithor/RL-Rust (paths below are under its `src/`), so client code reads like code written against the crate.
//
//     rlrust::TaxiEnv env(100);                                                        // env/taxi.rs:57
//     rlrust::TabularPolicy policy(0.05, 0.0);                                         // policy/tabular_policy.rs:15
//     rlrust::UniformEpsilonGreed selector(1.0, rlrust::Decay::sub(2e-5), 0.0);        // action_selection/uniform_epsilon_greed.rs:31
//     rlrust::OneStepAgent agent(policy, 0.95, selector, rlrust::qlearning);           // agent/one_step_agent.rs:16
//     auto [reward_history, episode_length, training_error] = agent.train(env, 100000, 10000);   // agent.rs:66-118
//
// One Agent object = `batch.n_agents` independent reference agents, each on its own Philox stream (n_agents = 1 is the
// reference's single trait object).  Differences forced by the device boundary: the epsilon-decay closure is
// Decay::sub(k) / Decay::mul(k); observations are dense indices (BlackJackEnv::obs_id gives the fxhash id);
// With ONE agent `training_error` is the reference's vector — one TD per training step (agent.rs:98,117); with a batch
// it is per episode (the sum of the episode's TDs), n_agents * n_episodes entries.
#pragma once
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "rlb.h"

namespace rlrust {

struct EnvNotReady : std::runtime_error {   // env.rs:16-17
    EnvNotReady() : std::runtime_error("EnvNotReady") {}
};
struct RlbError : std::runtime_error {
    rlb_status status;
    RlbError(rlb_status st, const std::string& what) : std::runtime_error(what), status(st) {}
};
inline void check(rlb_status st) {
    if (st == RLB_OK) return;
    if (st == RLB_ERR_ENV_NOT_READY) throw EnvNotReady();
    throw RlbError(st, rlb_last_error_string());
}

// agent.rs:17-45 — GetNextQValue, passed by name
enum GetNextQValue { sarsa = RLB_TARGET_SARSA, qlearning = RLB_TARGET_QLEARNING, expected_sarsa = RLB_TARGET_EXPECTED_SARSA };

struct Decay {   // stands for the `Rc<dyn Fn(f64) -> f64>` of uniform_epsilon_greed.rs:14
    rlb_decay_kind kind;
    double k;
    static Decay sub(double k) { return {RLB_DECAY_SUB, k}; }   // |a| a - k   (bin/taxi.rs:132)
    static Decay mul(double k) { return {RLB_DECAY_MUL, k}; }   // |a| a * k   (bin/frozen_lake_neural.rs:181)
};

struct Batch {   // the RNG injection contract and the device placement: no counterpart in the reference
    uint64_t n_agents = 1, seed = 0x5EED0001ull, first_agent_id = 0;
    rlb_real_kind real = RLB_REAL_F64;   // the reference's own arithmetic
    int device = 0;
};

class Engine {   // RAII over rlb_engine
   public:
    explicit Engine(const rlb_config& cfg) : cfg_(cfg) { check(rlb_engine_create(&cfg_, &e_)); }
    ~Engine() { rlb_engine_destroy(e_); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    rlb_engine* get() const { return e_; }
    const rlb_config& config() const { return cfg_; }

   private:
    rlb_config cfg_;
    rlb_engine* e_ = nullptr;
};

// ---------------------------------------------------------------------------------- Env<T, COUNT>  (env.rs:19-49)
// The cursor fix-up every render() of the reference does (e.g. env/taxi.rs:164-168): walk the newline offsets in order
// and push the cursor one to the right for each one at or before it (the cursor moves while they are walked).
inline size_t skip_newlines(const std::string& text, size_t pos) {
    for (size_t i = 0; i < text.size(); ++i)
        if (text[i] == '\n' && pos >= i) pos += 1;
    return pos;
}
inline std::string join_rows(const std::vector<std::string>& rows) {
    std::string out;
    for (size_t i = 0; i < rows.size(); ++i) { if (i) out += '\n'; out += rows[i]; }
    return out;
}

class Env {
   public:
    static constexpr size_t TRACK_LIMIT = 4096;   // render() follows step-level calls of engines up to this many agents
    virtual ~Env() = default;
    virtual size_t action_size() const = 0;                                  // env.rs:20-22
    virtual void describe(rlb_config& cfg) const = 0;
    virtual const char* get_action_label(size_t action) const = 0;           // env.rs:48
    std::vector<uint32_t> reset() {                                          // env.rs:23
        std::vector<uint32_t> obs(n());
        const std::vector<uint64_t> n0 = tracking() ? stream_positions() : std::vector<uint64_t>();
        check(rlb_env_reset(bound(), obs.data()));
        if (tracking()) {
            pos_.assign(obs.begin(), obs.end());
            nsteps_.assign(n(), 0);
            after_reset(n0, stream_positions());
        } else {
            pos_.clear();
        }
        return obs;
    }
    // env.rs:24 — throws EnvNotReady where the reference returns Err(EnvNotReady)
    std::tuple<std::vector<uint32_t>, std::vector<double>, std::vector<uint8_t>> step(const std::vector<uint32_t>& action) {
        std::vector<uint32_t> obs(n());
        std::vector<double> reward(n());
        std::vector<uint8_t> terminated(n());
        const bool track = !pos_.empty();
        const std::vector<uint64_t> n0 = track ? stream_positions() : std::vector<uint64_t>();
        check(rlb_env_step(bound(), action.data(), obs.data(), reward.data(), terminated.data(), nullptr));
        if (track) {
            for (size_t i = 0; i < n(); ++i) {   // a truncated step answers obs 0 WITHOUT moving (e.g. taxi.rs:148-151)
                if (nsteps_[i] >= step_limit()) continue;
                pos_[i] = obs[i];
                nsteps_[i] += 1;
            }
            after_step(action, n0, stream_positions());
        }
        return {obs, reward, terminated};
    }
    // env.rs:47 — of ONE of the batched envs, as the step-level reset()/step() calls since the last reset() left it
    std::string render(size_t agent = 0) const {
        if (pos_.empty()) throw std::logic_error("render() follows step-level reset()/step() calls: call reset() first");
        return render_at(pos_.at(agent), agent);
    }
    void bind(Engine* e) { engine_ = e; pos_.clear(); }
    Engine* engine() const { return engine_; }

   protected:
    virtual std::string render_at(uint32_t pos, size_t agent) const = 0;
    virtual uint64_t step_limit() const { return std::numeric_limits<uint64_t>::max(); }
    virtual void after_reset(const std::vector<uint64_t>&, const std::vector<uint64_t>&) {}
    virtual void after_step(const std::vector<uint32_t>&, const std::vector<uint64_t>&, const std::vector<uint64_t>&) {}
    std::vector<rlb_agent_state> states() const {
        std::vector<rlb_agent_state> st(n());
        check(rlb_get_agent_states(bound(), st.data()));
        return st;
    }
    size_t n() const { return engine_ ? (size_t)engine_->config().n_agents : 0; }

   private:
    bool tracking() const { return n() > 0 && n() <= TRACK_LIMIT; }
    std::vector<uint64_t> stream_positions() const {
        std::vector<uint64_t> out;
        for (const rlb_agent_state& s : states()) out.push_back(s.rng_n);
        return out;
    }
    rlb_engine* bound() const {
        if (!engine_) throw std::logic_error("env is not bound to an engine yet (train an agent on it first)");
        return engine_->get();
    }
    Engine* engine_ = nullptr;
    std::vector<uint32_t> pos_;
    std::vector<uint64_t> nsteps_;
};

class BlackJackEnv : public Env {   // env/blackjack.rs:30-188
   public:
    BlackJackEnv() = default;
    size_t action_size() const override { return 2; }
    void describe(rlb_config& c) const override { c.env_kind = RLB_ENV_BLACKJACK; }
    const char* get_action_label(size_t a) const override { static const char* k[] = {"HIT", "STICK"}; return k[a]; }   // :44
    static uint64_t obs_id(uint32_t dense) { return rlb_blackjack_obs_id(dense); }   // blackjack.rs:25-27
    static uint32_t dense_index(uint64_t id) { return rlb_blackjack_dense_index(id); }

   protected:
    // The engine keeps of a hand only what the rules read; the cards — which only render() shows — are re-derived from
    // the agent's Philox stream: an env call draws nothing but cards, so the words between the stream positions before
    // and after it, through rand's Uniform<u8>(1..11) (rlb_rng_card), are its cards (blackjack.rs:54-56,60-66,76).
    std::vector<uint32_t> cards_between(size_t agent, uint64_t n0, uint64_t n1) const {
        const rlb_config& c = engine()->config();
        std::vector<uint32_t> cards;
        uint64_t w = n0;
        while (w < n1) cards.push_back(rlb_rng_card(c.seed, c.first_agent_id + agent, &w));
        return cards;
    }
    void after_reset(const std::vector<uint64_t>& n0, const std::vector<uint64_t>& n1) override {
        player_.assign(n(), {});
        dealer_.assign(n(), {});
        for (size_t i = 0; i < n(); ++i) {
            const std::vector<uint32_t> c = cards_between(i, n0[i], n1[i]);
            player_[i] = {c.at(0), c.at(1)};
            dealer_[i] = {c.at(2), c.at(3)};
        }
    }
    void after_step(const std::vector<uint32_t>& action, const std::vector<uint64_t>& n0, const std::vector<uint64_t>& n1) override {
        for (size_t i = 0; i < n(); ++i) {
            std::vector<uint32_t>& hand = action[i] == 0 ? player_[i] : dealer_[i];   // HIT: the player; else the dealer to >= 17
            for (uint32_t c : cards_between(i, n0[i], n1[i])) hand.push_back(c);
        }
    }
    std::string render_at(uint32_t, size_t agent) const override {   // blackjack.rs:165-184
        std::string out = "Dealer: ";
        if (states().at(agent).env_ready) out += std::to_string(dealer_.at(agent).at(0));
        else for (uint32_t c : dealer_.at(agent)) out += std::to_string(c) + " ";
        out += " \nPlayer: ";
        for (uint32_t c : player_.at(agent)) out += std::to_string(c) + " ";
        return out;
    }

   private:
    std::vector<std::vector<uint32_t>> player_, dealer_;
};
class FrozenLakeEnv : public Env {   // env/frozen_lake.rs:12-134
   public:
    enum Map { MAP_4X4 = 0, MAP_8X8 = 1 };   // the crate's two constants (:23-28)
    FrozenLakeEnv(Map map, bool is_slippery, uint32_t max_steps)
        : rows_(map == MAP_4X4 ? std::vector<std::string>{"SFFF", "FHFH", "FFFH", "HFFG"}
                               : std::vector<std::string>{"SFFFFFFF", "FFFFFFFF", "FFFHFFFF", "FFFFFHFF", "FFFHFFFF", "FHHFFFHF", "FHFFHFHF", "FFFHFFFG"}),
          map_id_(map), slippery_(is_slippery), max_steps_(max_steps) {}
    // FrozenLakeEnv::new(map: &[&str], ..) (:48): any rows of S / F / H / G cells; every 'S' is a start cell (:54-66)
    FrozenLakeEnv(const std::vector<std::string>& map, bool is_slippery, uint32_t max_steps)
        : rows_(map), map_id_(RLB_MAP_CUSTOM), slippery_(is_slippery), max_steps_(max_steps) {
        if (rows_.empty()) throw std::invalid_argument("empty map");
        for (const std::string& r : rows_) {
            if (r.size() != rows_[0].size() || r.empty()) throw std::invalid_argument("map rows must be equally long");
            flat_ += r;
        }
    }
    size_t action_size() const override { return 4; }
    void describe(rlb_config& c) const override {
        c.env_kind = RLB_ENV_FROZEN_LAKE; c.map_id = map_id_; c.slippery = slippery_; c.max_steps = max_steps_;
        if (map_id_ == RLB_MAP_CUSTOM) { c.map_rows = (uint32_t)rows_.size(); c.map_cols = (uint32_t)rows_[0].size(); c.map = flat_.c_str(); }
    }
    const char* get_action_label(size_t a) const override { static const char* k[] = {"LEFT", "DOWN", "RIGHT", "UP"}; return k[a]; }   // :30

   protected:
    uint64_t step_limit() const override { return max_steps_; }
    std::string render_at(uint32_t pos, size_t) const override {   // frozen_lake.rs:136-149: every 'S' becomes 'F', then '@'
        std::string text = join_rows(rows_);
        for (char& ch : text) if (ch == 'S') ch = 'F';
        text[skip_newlines(text, pos)] = '@';
        return text;
    }

   private:
    std::vector<std::string> rows_;
    std::string flat_;
    int32_t map_id_; bool slippery_; uint32_t max_steps_;
};
class CliffWalkingEnv : public Env {   // env/cliff_walking.rs:6-89
   public:
    explicit CliffWalkingEnv(uint32_t max_steps) : max_steps_(max_steps) {}
    size_t action_size() const override { return 4; }
    void describe(rlb_config& c) const override { c.env_kind = RLB_ENV_CLIFF_WALKING; c.max_steps = max_steps_; }
    const char* get_action_label(size_t a) const override { static const char* k[] = {"LEFT", "DOWN", "RIGHT", "UP"}; return k[a]; }   // :19

   protected:
    uint64_t step_limit() const override { return max_steps_; }
    std::string render_at(uint32_t pos, size_t) const override {   // cliff_walking.rs:91-102: byte 39 (the start marker) becomes '_', then '@'
        std::string text = "____________\n____________\n____________\n@!!!!!!!!!!G";   // :20
        text[39] = '_';
        text[skip_newlines(text, pos)] = '@';
        return text;
    }

   private:
    uint32_t max_steps_;
};
class TaxiEnv : public Env {   // env/taxi.rs:10-159
   public:
    explicit TaxiEnv(uint32_t max_steps) : max_steps_(max_steps) {}
    size_t action_size() const override { return 6; }
    void describe(rlb_config& c) const override { c.env_kind = RLB_ENV_TAXI; c.max_steps = max_steps_; }
    const char* get_action_label(size_t a) const override {   // :31
        static const char* k[] = {"DOWN", "UP", "RIGHT", "LEFT", "PICKUP", "DROPOFF"};
        return k[a];
    }

   protected:
    uint64_t step_limit() const override { return max_steps_; }
    std::string render_at(uint32_t curr_obs, size_t) const override {   // taxi.rs:161-172: 'T' on the taxi's cell
        static const std::vector<std::string> map = {"+---------+", "|R: | : :G|", "| : | : : |", "| : : : : |", "| | : | : |", "|Y| : |B: |", "+---------+"};   // :21-29
        const size_t row = curr_obs / 100, col = (curr_obs / 20) % 5;   // decode, :44-55
        std::string text = join_rows(map);
        text[skip_newlines(text, 11 * (row + 1) + 2 * col + 1)] = 'T';   // from_2d_to_1d(11, row + 1, 2 * col + 1), utils.rs:45-47
        return text;
    }

   private:
    uint32_t max_steps_;
};

// ---------------------------------------------------------------------------------- Policy<T, COUNT>  (policy.rs:15-25)
struct TabularPolicy {   // policy/tabular_policy.rs:8-44
    double learning_rate, default_value;
    TabularPolicy(double lr, double dflt) : learning_rate(lr), default_value(dflt) {}
    virtual ~TabularPolicy() = default;
    virtual rlb_policy_kind kind() const { return RLB_POLICY_BASIC; }
};
struct DoubleTabularPolicy : TabularPolicy {   // policy/double_tabular_policy.rs:8-67
    using TabularPolicy::TabularPolicy;
    rlb_policy_kind kind() const override { return RLB_POLICY_DOUBLE; }
};

// ---------------------------------------------------------------------------------- ActionSelection<T, COUNT>
struct ActionSelection {   // action_selection.rs:10-15
    virtual ~ActionSelection() = default;
    virtual rlb_selector_kind kind() const = 0;
    virtual void describe(rlb_config& c) const = 0;
};
struct UniformEpsilonGreed : ActionSelection {   // action_selection/uniform_epsilon_greed.rs:8-80
    double epsilon; Decay decay; double final_epsilon;
    UniformEpsilonGreed(double eps, Decay d, double fin) : epsilon(eps), decay(d), final_epsilon(fin) {}
    rlb_selector_kind kind() const override { return RLB_SEL_EPS_GREEDY; }
    void describe(rlb_config& c) const override { c.initial_epsilon = epsilon; c.decay_kind = decay.kind; c.epsilon_decay = decay.k; c.final_epsilon = final_epsilon; }
};
struct UpperConfidenceBound : ActionSelection {   // action_selection/upper_confidence_bound.rs:9-68
    double confidence_level;
    explicit UpperConfidenceBound(double c) : confidence_level(c) {}
    rlb_selector_kind kind() const override { return RLB_SEL_UCB; }
    void describe(rlb_config& c) const override { c.confidence_level = confidence_level; }
};

// ---------------------------------------------------------------------------------- Agent<T, COUNT>  (agent.rs:47-164)
using TrainResult = std::tuple<std::vector<double>, std::vector<uint64_t>, std::vector<double>>;   // rewards, lengths, errors
using EvalResult = std::tuple<std::vector<double>, std::vector<uint64_t>>;

class Agent {
   public:
    virtual ~Agent() { delete engine_; }
    void set_future_q_value_func(GetNextQValue f) {                          // agent.rs:48
        target_ = f;
        if (engine_) check(rlb_agent_set_future_q_value_func(engine_->get(), f));
    }
    // agent.rs:50 — installs a fresh selector of that kind; call register_selector() before the first train() for
    // every selector whose parameters the engine must know (the bins build both up front, bin/taxi.rs:129-136)
    void set_action_selector(const ActionSelection& s) {
        s.describe(cfg_);
        cfg_.selector_kind = s.kind();
        if (engine_) check(rlb_agent_set_action_selector(engine_->get(), s.kind()));
    }
    void register_selector(const ActionSelection& s) { s.describe(cfg_); }
    std::vector<uint32_t> get_action(const std::vector<uint32_t>& obs) {     // agent.rs:52
        std::vector<uint32_t> a(obs.size());
        check(rlb_agent_get_action(need(), obs.data(), a.data()));
        return a;
    }
    // agent.rs:54-62; reward as f64, the returned temporal differences widened to f64
    std::vector<double> update(const std::vector<uint32_t>& curr_obs, const std::vector<uint32_t>& curr_action, const std::vector<double>& reward,
                               const std::vector<uint8_t>& terminated, const std::vector<uint32_t>& next_obs, const std::vector<uint32_t>& next_action) {
        std::vector<double> td(curr_obs.size());
        if (cfg_.real_kind == RLB_REAL_F64) {
            check(rlb_agent_update(need(), curr_obs.data(), curr_action.data(), reward.data(), terminated.data(), next_obs.data(), next_action.data(), td.data()));
        } else {
            std::vector<float> t32(curr_obs.size());
            check(rlb_agent_update(need(), curr_obs.data(), curr_action.data(), reward.data(), terminated.data(), next_obs.data(), next_action.data(), t32.data()));
            for (size_t i = 0; i < t32.size(); ++i) td[i] = t32[i];
        }
        return td;
    }
    void reset() { if (engine_) check(rlb_agent_reset(engine_->get())); }    // agent.rs:64

    // agent.rs:66-118.  [agent-major] vectors of n_agents * n_episodes entries (just n_episodes for one agent).
    TrainResult train(Env& env, uint64_t n_episodes, uint64_t eval_at) {
        if (eval_at == 0) throw std::domain_error("attempt to calculate the remainder with a divisor of zero");   // agent.rs:107
        bind(env);
        const uint64_t N = cfg_.n_agents;
        rlb_train_out out{};
        std::vector<rlb_episode_f64> e64;
        std::vector<rlb_episode_f32> e32;
        if (cfg_.real_kind == RLB_REAL_F64) { e64.resize(N * n_episodes); out.episodes = e64.data(); }
        else { e32.resize(N * n_episodes); out.episodes = e32.data(); }
        // one agent: the per-step TD stream; at most max_steps + 1 steps per episode (Blackjack: a hand holds 16 cards)
        std::vector<double> td64;
        std::vector<float> td32;
        uint64_t td_n = 0;
        if (N == 1) {
            const uint64_t cap = n_episodes * (cfg_.env_kind == RLB_ENV_BLACKJACK ? 32u : (uint64_t)cfg_.max_steps + 1u);
            if (cfg_.real_kind == RLB_REAL_F64) { td64.resize(cap); out.td_steps = td64.data(); }
            else { td32.resize(cap); out.td_steps = td32.data(); }
            out.td_capacity = cap; out.td_count = &td_n;
        }
        check(rlb_agent_train(engine_->get(), n_episodes, eval_at, &out));
        last_ = out;
        last_.td_steps = nullptr; last_.td_count = nullptr;
        TrainResult r;
        auto& [rew, len, err] = r;
        rew.resize(N * n_episodes); len.resize(N * n_episodes); err.resize(N == 1 ? td_n : N * n_episodes);
        if (N == 1) for (uint64_t k = 0; k < td_n; ++k) err[k] = cfg_.real_kind == RLB_REAL_F64 ? td64[k] : (double)td32[k];
        for (uint64_t ep = 0; ep < n_episodes; ++ep)
            for (uint64_t a = 0; a < N; ++a) {   // engine layout is [episode][agent]
                const uint64_t src = ep * N + a, dst = a * n_episodes + ep;
                if (cfg_.real_kind == RLB_REAL_F64) { rew[dst] = e64[src].ret; len[dst] = e64[src].length; if (N > 1) err[dst] = e64[src].td_sum; }
                else { rew[dst] = e32[src].ret; len[dst] = e32[src].length; if (N > 1) err[dst] = e32[src].td_sum; }
            }
        return r;
    }
    // agent.rs:120-141
    EvalResult evaluate(Env& env, uint64_t n_episodes) {
        bind(env);
        const uint64_t N = cfg_.n_agents;
        std::vector<rlb_episode_f64> e64;
        std::vector<rlb_episode_f32> e32;
        void* dst;
        if (cfg_.real_kind == RLB_REAL_F64) { e64.resize(N * n_episodes); dst = e64.data(); }
        else { e32.resize(N * n_episodes); dst = e32.data(); }
        uint64_t steps = 0;
        check(rlb_agent_evaluate(engine_->get(), n_episodes, dst, nullptr, &steps));
        EvalResult r;
        auto& [rew, len] = r;
        rew.resize(N * n_episodes); len.resize(N * n_episodes);
        for (uint64_t ep = 0; ep < n_episodes; ++ep)
            for (uint64_t a = 0; a < N; ++a) {
                const uint64_t src = ep * N + a, d = a * n_episodes + ep;
                if (cfg_.real_kind == RLB_REAL_F64) { rew[d] = e64[src].ret; len[d] = e64[src].length; }
                else { rew[d] = e32[src].ret; len[d] = e32[src].length; }
            }
        return r;
    }
    // agent.rs:143-163 — one episode through the step-level calls, printing what the reference prints: the env before
    // each step, the action's label (`{:?}` of a &str: quoted), the step's reward; at the end the last view, the
    // episode's reward and its length.  Needs an engine of ONE agent (the reference's single env).  Returns the lines.
    std::vector<std::string> example(Env& env, std::ostream& os = std::cout) {
        bind(env);
        if (cfg_.n_agents != 1) throw std::logic_error("example() shows the reference's single env: use an engine of one agent");
        auto f64_debug = [](double x) {   // `{:?}` of the f64 values an episode produces (integers: "-1.0", "20.0")
            char buf[64];
            if (x == (double)(long long)x && x > -1e15 && x < 1e15) std::snprintf(buf, sizeof buf, "%lld.0", (long long)x);
            else std::snprintf(buf, sizeof buf, "%.17g", x);
            return std::string(buf);
        };
        std::vector<std::string> lines;
        double epi_reward = 0.0;
        std::vector<uint32_t> curr_action = get_action(env.reset());
        int steps = 0;
        for (;;) {
            steps += 1;
            lines.push_back(env.render());
            auto [next_obs, reward, terminated] = env.step(curr_action);
            const std::vector<uint32_t> next_action = get_action(next_obs);   // also on the terminal observation (:153)
            lines.push_back(std::string("\"") + env.get_action_label(curr_action[0]) + "\"");
            lines.push_back("step reward " + f64_debug(reward[0]));
            curr_action = next_action;
            epi_reward += reward[0];
            if (terminated[0]) {
                lines.push_back(env.render());
                lines.push_back("episode reward " + f64_debug(epi_reward));
                lines.push_back("terminated with " + std::to_string(steps) + " steps");
                break;
            }
        }
        for (const std::string& l : lines) os << l << "\n";
        return lines;
    }
    const rlb_train_out& last_train() const { return last_; }   // step totals, kernel time of the last train()
    Engine* engine() const { return engine_; }

   protected:
    Agent(const TabularPolicy& policy, double discount_factor, const ActionSelection& selector, double lambda_factor,
          GetNextQValue f, rlb_agent_kind kind, const Batch& b)
        : target_(f) {
        cfg_ = rlb_config{};
        cfg_.struct_size = sizeof(rlb_config);
        cfg_.policy_kind = policy.kind();
        cfg_.learning_rate = policy.learning_rate;
        cfg_.default_value = policy.default_value;
        cfg_.discount_factor = discount_factor;
        cfg_.lambda_factor = lambda_factor;
        cfg_.agent_kind = kind;
        cfg_.confidence_level = 0.5;   // bin/taxi.rs:54 default until a UCB selector says otherwise
        cfg_.initial_epsilon = 1.0;
        selector.describe(cfg_);
        cfg_.selector_kind = selector.kind();
        cfg_.real_kind = b.real; cfg_.device = b.device; cfg_.seed = b.seed; cfg_.n_agents = b.n_agents; cfg_.first_agent_id = b.first_agent_id;
        cfg_.max_steps = 100;
    }

   private:
    friend class InternalModelAgent;
    void bind(Env& env) {
        if (engine_) {
            if (env.engine() != engine_) throw std::logic_error("agent is already bound to another env");
            return;
        }
        env.describe(cfg_);
        cfg_.target_kind = target_;
        engine_ = new Engine(cfg_);
        env.bind(engine_);
    }
    rlb_engine* need() const {
        if (!engine_) throw std::logic_error("agent is not bound yet: train or evaluate on an env first");
        return engine_->get();
    }
    rlb_config cfg_;
    GetNextQValue target_;
    Engine* engine_ = nullptr;
    rlb_train_out last_{};
};

class OneStepAgent : public Agent {   // agent/one_step_agent.rs:7-86
   public:
    OneStepAgent(const TabularPolicy& policy, double discount_factor, const ActionSelection& action_selection, GetNextQValue f,
                 const Batch& batch = Batch())
        : Agent(policy, discount_factor, action_selection, 0.0, f, RLB_AGENT_ONE_STEP, batch) {}
};
class ElegibilityTracesAgent : public Agent {   // agent/elegibility_traces_agent.rs:8-104
   public:
    ElegibilityTracesAgent(const TabularPolicy& policy, double discount_factor, const ActionSelection& action_selection,
                           double lambda_factor, GetNextQValue f, const Batch& batch = Batch())
        : Agent(policy, discount_factor, action_selection, lambda_factor, f, RLB_AGENT_TRACES, batch) {}
};

// ---------------------------------------------------------------------------------- Model<T, COUNT>  (model.rs:12-16)
class RandomModel {   // model/random_model.rs:9-45 — lives on the device, next to the agent its InternalModelAgent borrows
   public:
    struct Info { std::vector<uint32_t> obs, action, next_obs; std::vector<double> reward; };   // (state, action, next_state, reward)
    Info get_info() {                                                                      // :27-35
        const size_t n = agents();
        Info r{std::vector<uint32_t>(n), std::vector<uint32_t>(n), std::vector<uint32_t>(n), std::vector<double>(n)};
        check(rlb_model_get_info(need(), r.obs.data(), r.action.data(), r.next_obs.data(), r.reward.data()));
        return r;
    }
    void add_info(const std::vector<uint32_t>& obs, const std::vector<uint32_t>& action, const std::vector<double>& reward,
                  const std::vector<uint32_t>& next_obs) {                                 // :37-41
        check(rlb_model_add_info(need(), obs.data(), action.data(), reward.data(), next_obs.data()));
    }
    void reset() { check(rlb_model_reset(need())); }                                       // :43-45
    // remembered transitions of every agent, in insertion order: len [n_agents], entries [n_agents][capacity]
    std::pair<std::vector<uint32_t>, std::vector<rlb_model_entry>> entries() {
        const size_t n = agents(), cap = rlb_model_capacity(need());
        std::vector<uint32_t> len(n);
        std::vector<rlb_model_entry> ent(n * cap);
        check(rlb_download_model(need(), len.data(), ent.data()));
        return {len, ent};
    }

   private:
    friend class InternalModelAgent;
    rlb_engine* need() const {
        if (!engine_) throw std::logic_error("the model is bound when its InternalModelAgent first sees an env");
        return engine_->get();
    }
    size_t agents() const { return n_agents_; }
    Engine* engine_ = nullptr;
    size_t n_agents_ = 0;
};

// agent/internal_model_agent.rs:9-85 — Dyna.  Borrows `agent` (which keeps what it has learned) and `model` for its
// lifetime, like the `&'a mut dyn Agent` of the reference; planning_length must be > 0.
class InternalModelAgent {
   public:
    InternalModelAgent(Agent& agent, RandomModel& model, uint32_t planning_length) : agent_(agent), model_(model), planning_(planning_length) {
        if (!planning_length) throw std::invalid_argument("planning_length must be > 0");
        if (agent_.engine_) attach();
    }
    ~InternalModelAgent() {   // end of the borrow: the agent carries on without a model
        if (agent_.engine_ && model_.engine_ == agent_.engine_) rlb_agent_set_model(agent_.engine_->get(), 0);
        model_.engine_ = nullptr;
    }
    InternalModelAgent(const InternalModelAgent&) = delete;
    InternalModelAgent& operator=(const InternalModelAgent&) = delete;
    void set_future_q_value_func(GetNextQValue f) { agent_.set_future_q_value_func(f); }          // :34-36
    void set_action_selector(const ActionSelection& s) { agent_.set_action_selector(s); }         // :38-40
    std::vector<uint32_t> get_action(const std::vector<uint32_t>& obs) { return agent_.get_action(obs); }   // :42-44
    std::vector<double> update(const std::vector<uint32_t>& curr_obs, const std::vector<uint32_t>& curr_action, const std::vector<double>& reward,
                               const std::vector<uint8_t>& terminated, const std::vector<uint32_t>& next_obs,
                               const std::vector<uint32_t>& next_action) {                        // :46-79
        attach();
        return agent_.update(curr_obs, curr_action, reward, terminated, next_obs, next_action);
    }
    void reset() { agent_.reset(); }                                                              // :81-84 (the engine empties the attached model too)
    TrainResult train(Env& env, uint64_t n_episodes, uint64_t eval_at) {
        agent_.bind(env);
        attach();
        return agent_.train(env, n_episodes, eval_at);
    }
    EvalResult evaluate(Env& env, uint64_t n_episodes) {
        agent_.bind(env);
        attach();
        return agent_.evaluate(env, n_episodes);
    }

   private:
    void attach() {
        if (!agent_.engine_) throw std::logic_error("agent is not bound yet: train or evaluate on an env first");
        if (model_.engine_ == agent_.engine_) return;
        check(rlb_agent_set_model(agent_.engine_->get(), planning_));
        model_.engine_ = agent_.engine_;
        model_.n_agents_ = agent_.cfg_.n_agents;
    }
    Agent& agent_;
    RandomModel& model_;
    uint32_t planning_;
};

}   // namespace rlrust
