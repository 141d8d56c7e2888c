// oracle/oracle_capi.cpp — C entry points over oracle.hpp for ctypes (tests/, smoke(),
// bench.py's cpu_baseline and `--impl reference` legs).  TEST INFRASTRUCTURE ONLY: the
// product library (rl-rust_b200/librlb.so) never links or calls this.
//
// Build: make -C oracle   (g++ -O2 -ffp-contract=off -shared -fPIC -pthread)
#include "oracle.hpp"

#include <atomic>
#include <chrono>
#include <thread>

using namespace oracle;

extern "C" {

struct OracleConfig {
    int32_t env_kind;        // 0 blackjack, 1 frozen_lake, 2 cliff_walking, 3 taxi
    int32_t map_id;          // frozen_lake: 0 = 4x4, 1 = 8x8
    int32_t slippery;
    uint32_t max_steps;
    int32_t policy_kind;     // 0 basic, 1 double
    int32_t selector_kind;   // 0 eps-greedy, 1 ucb
    int32_t target_kind;     // 0 sarsa, 1 qlearning, 2 expected_sarsa
    int32_t agent_kind;      // 0 one-step, 1 eligibility traces
    int32_t real_kind;       // 0 f32, 1 f64
    int32_t decay_kind;      // 0 eps - param, 1 eps * param
    double lr, gamma, lambda, eps0, eps_decay, eps_final, ucb_c, default_q;
    uint64_t seed;
    uint32_t planning_steps; // > 0: the agent is wrapped in InternalModelAgent(RandomModel, planning_steps)
    uint32_t pad;
    // frozen_lake with map_id == 2: FrozenLakeEnv::new(map, ..) on the caller's own rows (frozen_lake.rs:48), joined
    // without separators, map_rows * map_cols cells
    uint32_t map_rows, map_cols;
    const char* map;
};

struct OracleState {
    double epsilon;
    uint64_t ucb_t;
    uint64_t rng_n;
    uint64_t eval_steps;
    int32_t policy_flag;
    int32_t pad;
};

}   // extern "C"

namespace {

struct SessionBase {
    virtual ~SessionBase() {}
    virtual int train(u64 ep_begin, u64 ep_end, u64 eval_at, double* ret, u64* len, double* tdsum, double* tdabs) = 0;
    virtual int evaluate(u64 n, double* ret, u64* len) = 0;
    virtual void agent_reset() = 0;
    virtual void set_target(int kind) = 0;
    virtual void set_selector(int kind) = 0;
    virtual void set_agent_kind(int kind) = 0;
    virtual void export_tables(double* q, u64* counts, OracleState* st) = 0;
    virtual u32 n_states() = 0;
    virtual u32 n_actions() = 0;
    virtual u32 n_tables() = 0;
    virtual u32 env_reset() = 0;
    virtual int env_step(u32 action, u32* obs, double* reward, int* term) = 0;
    virtual u32 get_action(u32 dense_obs) = 0;
    virtual double update(u32 s, u32 a, double r, int term, u32 s2, u32 a2) = 0;
    virtual void record(bool on) = 0;
    virtual void set_planning(u32 steps) = 0;
    virtual u64 model_len() = 0;
    virtual void model_copy(u32* s, u32* a, u32* s2, double* r) = 0;
    virtual void model_add_info(u32 s, u32 a, double r, u32 s2) = 0;
    virtual int model_get_info(u32* s, u32* a, u32* s2, double* r) = 0;
    virtual void model_reset() = 0;
    std::vector<TrajRecord> traj;
    std::vector<double> training_error;   // widened copy of the last train() call's per-step TD
    u64 train_steps = 0;
};

template <int A, class Real>
struct Session : SessionBase {
    OracleConfig cfg;
    Stream rng;
    std::unique_ptr<Env<A>> env;
    std::unique_ptr<Agent<A, Real>> inner;            // the wrapped agent when a model is attached
    std::unique_ptr<Agent<A, Real>> agent;
    InternalModelAgent<A, Real>* dyna = nullptr;
    Policy<A, Real>* policy_raw = nullptr;            // borrowed views for export
    TabularPolicy<A, Real>* basic = nullptr;
    DoubleTabularPolicy<A, Real>* dbl = nullptr;
    ActionSelection<A, Real>* selector_raw = nullptr;
    std::vector<u64> id_of_dense;                      // dense index -> observation id

    static GetNextQValue<A, Real> target_fn(int kind) {
        switch (kind) {
            case TARGET_SARSA: return &sarsa<A, Real>;
            case TARGET_QLEARNING: return &qlearning<A, Real>;
            default: return &expected_sarsa<A, Real>;
        }
    }
    std::unique_ptr<ActionSelection<A, Real>> make_selector(int kind) {
        if (kind == 0)
            return std::unique_ptr<ActionSelection<A, Real>>(new UniformEpsilonGreed<A, Real>(
                cfg.eps0, cfg.decay_kind, cfg.eps_decay, cfg.eps_final, &rng));
        return std::unique_ptr<ActionSelection<A, Real>>(new UpperConfidenceBound<A, Real>(cfg.ucb_c));
    }
    Session(const OracleConfig& c, u64 agent_id, Env<A>* e_) : cfg(c), rng(c.seed, agent_id, 0), env(e_) {}
    // a second agent object over the same env and RNG stream (bin/taxi.rs:138-156)
    void set_agent_kind(int kind) override {
        cfg.agent_kind = kind;
        basic = nullptr; dbl = nullptr;
        build_agent();
    }
    void finish_init() {
        build_agent();
        u32 S = env->n_states();
        id_of_dense.resize(S);
        if (cfg.env_kind == 0) id_of_dense = BlackJackEnv::id_table();
        else for (u32 i = 0; i < S; ++i) id_of_dense[i] = i;
    }
    void build_agent() {
        std::unique_ptr<Policy<A, Real>> pol;
        if (cfg.policy_kind == 0) { basic = new TabularPolicy<A, Real>((Real)cfg.lr, (Real)cfg.default_q); pol.reset(basic); }
        else { dbl = new DoubleTabularPolicy<A, Real>((Real)cfg.lr, (Real)cfg.default_q); pol.reset(dbl); }
        policy_raw = pol.get();
        auto sel = make_selector(cfg.selector_kind);
        selector_raw = sel.get();
        if (cfg.agent_kind == 0)
            agent.reset(new OneStepAgent<A, Real>(std::move(pol), (Real)cfg.gamma, std::move(sel), target_fn(cfg.target_kind)));
        else
            agent.reset(new ElegibilityTracesAgent<A, Real>(std::move(pol), (Real)cfg.gamma, std::move(sel),
                                                            (Real)cfg.lambda, target_fn(cfg.target_kind)));
        dyna = nullptr;
        inner.reset();
        if (cfg.planning_steps) wrap_with_model();
    }
    // InternalModelAgent::new(Box::new(RefCell::new(&mut other)), EnumModel::from(RandomModel::default()), n)
    // (bin/cliffwalking_model.rs:150-156): the wrapped agent keeps its tables.
    void wrap_with_model() {
        inner = std::move(agent);
        dyna = new InternalModelAgent<A, Real>(inner.get(), RandomModel<A, Real>(&rng), cfg.planning_steps);
        agent.reset(dyna);
    }
    void set_planning(u32 steps) override {
        std::vector<TrajRecord>* tap = agent->recorder;
        u64 ev = agent->eval_steps;
        if (dyna) { agent = std::move(inner); dyna = nullptr; }   // unwrap (drops the model)
        cfg.planning_steps = steps;
        if (steps) wrap_with_model();
        agent->recorder = tap;
        agent->eval_steps = ev;
    }
    u64 model_len() override { return dyna ? dyna->model.values.size() : 0; }
    void model_copy(u32* s, u32* a, u32* s2, double* r) override {
        if (!dyna) return;
        size_t i = 0;
        for (auto& v : dyna->model.values) {
            s[i] = env->dense_index(v.obs); a[i] = (u32)v.action; s2[i] = env->dense_index(v.next_obs); r[i] = (double)v.reward;
            ++i;
        }
    }
    void model_add_info(u32 s, u32 a, double r, u32 s2) override { if (dyna) dyna->model.add_info(id_of_dense[s], a, (Real)r, id_of_dense[s2]); }
    int model_get_info(u32* s, u32* a, u32* s2, double* r) override {
        if (!dyna || dyna->model.values.empty()) return 1;   // the reference panics on an empty range
        auto v = dyna->model.get_info();
        *s = env->dense_index(v.obs); *a = (u32)v.action; *s2 = env->dense_index(v.next_obs); *r = (double)v.reward;
        return 0;
    }
    void model_reset() override { if (dyna) dyna->model.reset(); }
    int train(u64 ep_begin, u64 ep_end, u64 eval_at, double* ret, u64* len, double* tdsum, double* tdabs) override {
        std::vector<Real> r, te; std::vector<u64> l;
        if (eval_at == 0) return 2;   // reference: division by zero panic (agent.rs:107)
        bool ok = agent->train_range(*env, ep_begin, ep_end, eval_at, r, l, te);
        if (!ok) return 1;
        training_error.assign(te.begin(), te.end());
        size_t pos = 0;
        for (size_t e = 0; e < l.size(); ++e) {
            Real s = (Real)0, sa = (Real)0;
            for (u64 k = 0; k < l[e]; ++k, ++pos) { s += te[pos]; sa += (te[pos] < 0 ? -te[pos] : te[pos]); }
            if (ret) ret[e] = (double)r[e];
            if (len) len[e] = l[e];
            if (tdsum) tdsum[e] = (double)s;
            if (tdabs) tdabs[e] = (double)sa;
            train_steps += l[e];
        }
        return 0;
    }
    int evaluate(u64 n, double* ret, u64* len) override {
        std::vector<Real> r; std::vector<u64> l;
        if (!agent->evaluate(*env, n, r, l)) return 1;
        for (size_t e = 0; e < l.size(); ++e) { if (ret) ret[e] = (double)r[e]; if (len) len[e] = l[e]; }
        return 0;
    }
    void agent_reset() override { agent->reset(); }
    void set_target(int kind) override { cfg.target_kind = kind; agent->set_future_q_value_func(target_fn(kind)); }
    void set_selector(int kind) override {
        cfg.selector_kind = kind;
        auto sel = make_selector(kind);
        selector_raw = sel.get();
        agent->set_action_selector(std::move(sel));
    }
    u32 n_states() override { return env->n_states(); }
    u32 n_actions() override { return A; }
    u32 n_tables() override { return cfg.policy_kind == 0 ? 1 : 2; }
    void export_tables(double* q, u64* counts, OracleState* st) override {
        u32 S = env->n_states();
        auto dump = [&](FxMap<std::array<Real, A>>& m, const std::array<Real, A>& dflt, double* dst) {
            for (u32 s = 0; s < S; ++s) {
                const auto* row = m.get(id_of_dense[s]);
                for (int a = 0; a < A; ++a) dst[(size_t)s * A + a] = (double)(row ? (*row)[a] : dflt[a]);
            }
        };
        if (q) {
            if (basic) dump(basic->policy, basic->dflt, q);
            else { dump(dbl->alpha_policy, dbl->dflt, q); dump(dbl->beta_policy, dbl->dflt, q + (size_t)S * A); }
        }
        auto* ucb = dynamic_cast<UpperConfidenceBound<A, Real>*>(selector_raw);
        auto* eg = dynamic_cast<UniformEpsilonGreed<A, Real>*>(selector_raw);
        if (counts) {
            for (u32 s = 0; s < S; ++s) {
                const std::array<u64, A>* row = ucb ? ucb->action_counter.get(id_of_dense[s]) : nullptr;
                for (int a = 0; a < A; ++a) counts[(size_t)s * A + a] = row ? (*row)[a] : 0;
            }
        }
        if (st) {
            st->epsilon = eg ? eg->epsilon : 0.0;
            st->ucb_t = ucb ? ucb->t : 1;
            st->rng_n = rng.n;
            st->eval_steps = agent->eval_steps;
            st->policy_flag = dbl ? (dbl->policy_flag ? 1 : 0) : 1;
            st->pad = 0;
        }
    }
    u32 env_reset() override { return env->dense_index(env->reset()); }
    int env_step(u32 action, u32* obs, double* reward, int* term) override {
        StepResult sr;
        if (!env->step(action, sr)) return 1;
        *obs = env->dense_index(sr.obs); *reward = sr.reward; *term = sr.terminated ? 1 : 0;
        return 0;
    }
    u32 get_action(u32 dense_obs) override { return (u32)agent->get_action(id_of_dense[dense_obs]); }
    double update(u32 s, u32 a, double r, int term, u32 s2, u32 a2) override {
        return (double)agent->update(id_of_dense[s], a, (Real)r, term != 0, id_of_dense[s2], a2);
    }
    void record(bool on) override { agent->recorder = on ? &traj : nullptr; }
};

template <class Real>
SessionBase* make_session(const OracleConfig& c, u64 agent_id) {
    switch (c.env_kind) {
        case 0: {
            auto* s = new Session<2, Real>(c, agent_id, nullptr);
            s->env.reset(new BlackJackEnv(&s->rng));   // new() deals 4 cards from the agent's stream
            s->finish_init();
            return s;
        }
        case 1: {
            auto* s = new Session<4, Real>(c, agent_id, nullptr);
            std::vector<std::string> map = c.map_id == 0 ? FrozenLakeEnv::map_4x4() : FrozenLakeEnv::map_8x8();
            if (c.map_id == 2) {
                map.clear();
                for (uint32_t r = 0; r < c.map_rows; ++r) map.emplace_back(c.map + (size_t)r * c.map_cols, c.map_cols);
            }
            s->env.reset(new FrozenLakeEnv(map, c.slippery != 0, c.max_steps, &s->rng));
            s->finish_init();
            return s;
        }
        case 2: {
            auto* s = new Session<4, Real>(c, agent_id, nullptr);
            s->env.reset(new CliffWalkingEnv(c.max_steps));
            s->finish_init();
            return s;
        }
        case 3: {
            auto* s = new Session<6, Real>(c, agent_id, nullptr);
            s->env.reset(new TaxiEnv(c.max_steps, &s->rng));
            s->finish_init();
            return s;
        }
    }
    return nullptr;
}

}   // namespace

extern "C" {

void* oracle_create(const OracleConfig* cfg, uint64_t agent_id) {
    return cfg->real_kind == 0 ? (void*)make_session<float>(*cfg, agent_id) : (void*)make_session<double>(*cfg, agent_id);
}
void oracle_destroy(void* h) { delete (SessionBase*)h; }
uint32_t oracle_n_states(void* h) { return ((SessionBase*)h)->n_states(); }
uint32_t oracle_n_actions(void* h) { return ((SessionBase*)h)->n_actions(); }
uint32_t oracle_n_tables(void* h) { return ((SessionBase*)h)->n_tables(); }

// Agent::train (agent.rs:66-118), episodes [ep_begin, ep_end).  0 ok, 1 EnvNotReady, 2 eval_at == 0.
int oracle_train(void* h, uint64_t ep_begin, uint64_t ep_end, uint64_t eval_at, double* ret, uint64_t* len,
                 double* tdsum, double* tdabs) {
    return ((SessionBase*)h)->train(ep_begin, ep_end, eval_at, ret, len, tdsum, tdabs);
}
// Agent::evaluate (agent.rs:120-141)
int oracle_evaluate(void* h, uint64_t n, double* ret, uint64_t* len) { return ((SessionBase*)h)->evaluate(n, ret, len); }
void oracle_agent_reset(void* h) { ((SessionBase*)h)->agent_reset(); }
void oracle_set_target(void* h, int kind) { ((SessionBase*)h)->set_target(kind); }
void oracle_set_selector(void* h, int kind) { ((SessionBase*)h)->set_selector(kind); }
void oracle_set_agent_kind(void* h, int kind) { ((SessionBase*)h)->set_agent_kind(kind); }
void oracle_export(void* h, double* q, uint64_t* counts, OracleState* st) { ((SessionBase*)h)->export_tables(q, counts, st); }
uint64_t oracle_training_error_len(void* h) { return ((SessionBase*)h)->training_error.size(); }
void oracle_training_error_copy(void* h, double* out) {
    auto& v = ((SessionBase*)h)->training_error;
    std::memcpy(out, v.data(), v.size() * sizeof(double));
}
void oracle_record(void* h, int on) { ((SessionBase*)h)->record(on != 0); }
uint64_t oracle_traj_len(void* h) { return ((SessionBase*)h)->traj.size(); }
void oracle_traj_copy(void* h, TrajRecord* out) {
    auto& v = ((SessionBase*)h)->traj;
    std::memcpy((void*)out, v.data(), v.size() * sizeof(TrajRecord));
}
void oracle_traj_clear(void* h) { ((SessionBase*)h)->traj.clear(); }

// step-level trait methods
uint32_t oracle_env_reset(void* h) { return ((SessionBase*)h)->env_reset(); }
int oracle_env_step(void* h, uint32_t action, uint32_t* obs, double* reward, int* term) {
    return ((SessionBase*)h)->env_step(action, obs, reward, term);
}
uint32_t oracle_get_action(void* h, uint32_t dense_obs) { return ((SessionBase*)h)->get_action(dense_obs); }
double oracle_update(void* h, uint32_t s, uint32_t a, double r, int term, uint32_t s2, uint32_t a2) {
    return ((SessionBase*)h)->update(s, a, r, term, s2, a2);
}

// Model<T,COUNT> (model.rs:12-16) of the attached RandomModel, and (un)wrapping the agent
void oracle_set_planning(void* h, uint32_t steps) { ((SessionBase*)h)->set_planning(steps); }
uint64_t oracle_model_len(void* h) { return ((SessionBase*)h)->model_len(); }
void oracle_model_copy(void* h, uint32_t* s, uint32_t* a, uint32_t* s2, double* r) { ((SessionBase*)h)->model_copy(s, a, s2, r); }
void oracle_model_add_info(void* h, uint32_t s, uint32_t a, double r, uint32_t s2) { ((SessionBase*)h)->model_add_info(s, a, r, s2); }
int oracle_model_get_info(void* h, uint32_t* s, uint32_t* a, uint32_t* s2, double* r) { return ((SessionBase*)h)->model_get_info(s, a, s2, r); }
void oracle_model_reset(void* h) { ((SessionBase*)h)->model_reset(); }

// ------------------------------------------------------------------ batch runner
// Runs agents [first_agent, first_agent+count): create, train [0,n_episodes) with eval_at,
// export.  Any output pointer may be null.  Layouts: ret/len/tdsum/tdabs [count][n_episodes];
// q [count][tables][S][A]; counts [count][S][A]; states [count].  Returns train seconds
// (wall, create/export excluded) in *seconds, total steps in *train_steps / *eval_steps.
int oracle_batch_train(const OracleConfig* cfg, uint64_t first_agent, uint64_t count, uint64_t n_episodes,
                       uint64_t eval_at, int n_threads, double* ret, uint64_t* len, double* tdsum, double* tdabs,
                       double* q, uint64_t* counts, OracleState* states, double* seconds, uint64_t* train_steps,
                       uint64_t* eval_steps) {
    if (n_threads < 1) n_threads = 1;
    std::vector<SessionBase*> sessions(count);
    for (u64 i = 0; i < count; ++i) sessions[i] = (SessionBase*)oracle_create(cfg, first_agent + i);
    std::atomic<u64> next(0);
    std::atomic<int> rc(0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        for (;;) {
            u64 i = next.fetch_add(1);
            if (i >= count) break;
            int r = sessions[i]->train(0, n_episodes, eval_at, ret ? ret + i * n_episodes : nullptr,
                                       len ? len + i * n_episodes : nullptr, tdsum ? tdsum + i * n_episodes : nullptr,
                                       tdabs ? tdabs + i * n_episodes : nullptr);
            if (r) rc = r;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    u64 ts = 0, es = 0;
    for (u64 i = 0; i < count; ++i) {
        SessionBase* s = sessions[i];
        size_t SA = (size_t)s->n_states() * s->n_actions();
        OracleState st;
        s->export_tables(q ? q + i * SA * s->n_tables() : nullptr, counts ? counts + i * SA : nullptr, &st);
        if (states) states[i] = st;
        ts += s->train_steps;
        es += st.eval_steps;
        delete s;
    }
    if (train_steps) *train_steps = ts;
    if (eval_steps) *eval_steps = es;
    return rc;
}

// ------------------------------------------------------------------ unit-test hooks
void oracle_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
void oracle_stream_words(uint64_t seed, uint64_t agent, uint64_t start_n, uint64_t count, uint32_t* out) {
    Stream s(seed, agent, start_n);
    for (u64 i = 0; i < count; ++i) out[i] = s.next_u32();
}
// draws `count` values of one sampler kind from the stream starting at word start_n:
// kind 0 uniform_f64 (out as double), 1 uniform_usize(range), 2 card, 3 gen_range(0..range).  Returns words consumed.
uint64_t oracle_sample(uint64_t seed, uint64_t agent, uint64_t start_n, int kind, uint64_t range, uint64_t count, double* out) {
    Stream s(seed, agent, start_n);
    for (u64 i = 0; i < count; ++i) {
        if (kind == 0) out[i] = uniform_f64(s);
        else if (kind == 1) out[i] = (double)uniform_usize(s, range);
        else if (kind == 3) out[i] = (double)gen_range_usize(s, range);
        else out[i] = (double)uniform_card(s);
    }
    return s.n - start_n;
}
uint64_t oracle_fxhash_blackjack(uint32_t p, uint32_t d, int ace) { return fxhash_blackjack((u8)p, (u8)d, ace != 0); }
uint32_t oracle_blackjack_dense(uint32_t p, uint32_t d, int ace) { return BlackJackEnv::dense_of((u8)p, (u8)d, ace != 0); }
double oracle_log(double x) { return portable_log(x); }
uint64_t oracle_categorical_sample(const double* probs, uint64_t len, double random) { return categorical_sample(probs, len, random); }
uint64_t oracle_argmax(const double* v, uint64_t len) { return argmax<double>(v, len); }

// transition tables as the env constructors build them: out_s/out_r/out_t [S][A] (Taxi, Cliff)
void oracle_taxi_table(uint32_t* out_s, double* out_r, uint8_t* out_t, double* init_distrib) {
    Stream dummy;
    TaxiEnv env(100, &dummy);
    for (int s = 0; s < 500; ++s) for (int a = 0; a < 6; ++a) {
        out_s[s * 6 + a] = (u32)env.obs[s][a].s; out_r[s * 6 + a] = env.obs[s][a].r; out_t[s * 6 + a] = env.obs[s][a].t;
    }
    if (init_distrib) for (int s = 0; s < 500; ++s) init_distrib[s] = env.initial_state_distrib[s];
}
void oracle_cliff_table(uint32_t* out_s, double* out_r, uint8_t* out_t) {
    CliffWalkingEnv env(100);
    for (int s = 0; s < 48; ++s) for (int a = 0; a < 4; ++a) {
        out_s[s * 4 + a] = (u32)env.obs[s][a].s; out_r[s * 4 + a] = env.obs[s][a].r; out_t[s * 4 + a] = env.obs[s][a].t;
    }
}
// FrozenLake: [S][4][3] of (p, s', r, t)
void oracle_frozen_lake_table(int map_id, int slippery, double* out_p, uint32_t* out_s, double* out_r, uint8_t* out_t) {
    Stream dummy;
    FrozenLakeEnv env(map_id == 0 ? FrozenLakeEnv::map_4x4() : FrozenLakeEnv::map_8x8(), slippery != 0, 100, &dummy);
    size_t S = env.probs.size();
    for (size_t s = 0; s < S; ++s) for (int a = 0; a < 4; ++a) for (int i = 0; i < 3; ++i) {
        size_t o = (s * 4 + a) * 3 + i;
        out_p[o] = env.probs[s][a][i].p; out_s[o] = (u32)env.probs[s][a][i].s;
        out_r[o] = env.probs[s][a][i].r; out_t[o] = env.probs[s][a][i].t;
    }
}

}   // extern "C"
