// tests/cpp/mirror_main.cpp — client code written against include/rlb.hpp the way it would be written against the
// reference crate; prints what it gets so tests/test_cpp_mirror.py can compare with the oracle.
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <sstream>

#include "rlb.hpp"

using namespace rlrust;

static void dump(const char* tag, const std::vector<double>& v) {
    std::printf("%s", tag);
    for (double x : v) { uint64_t b; std::memcpy(&b, &x, 8); std::printf(" %016" PRIx64, b); }
    std::printf("\n");
}
static void dump(const char* tag, const std::vector<uint64_t>& v) {
    std::printf("%s", tag);
    for (uint64_t x : v) std::printf(" %" PRIu64, x);
    std::printf("\n");
}

int main(int argc, char** argv) {
    const bool probe = argc > 1 && std::strcmp(argv[1], "probe") == 0;
    try {
        // --- bin/taxi.rs in miniature: n_episodes = 60, three agents on streams 0..2 of seed 0xC0DE
        const uint64_t n_episodes = 60;
        const double epsilon_decay = 1.0 / (0.5 * (double)n_episodes);                     // bin/taxi.rs:78
        TaxiEnv env(100);
        TabularPolicy policy(0.05, 0.0);
        UniformEpsilonGreed eg(1.0, Decay::sub(epsilon_decay), 0.0);
        UpperConfidenceBound ucb(0.5);
        Batch batch;
        batch.n_agents = 3; batch.seed = 0xC0DE;
        OneStepAgent agent(policy, 0.95, eg, qlearning, batch);
        agent.register_selector(ucb);
        if (probe) {   // no GPU: the engine must refuse loudly
            try { agent.train(env, 1, 1); } catch (const RlbError& e) { std::printf("probe status=%d %s\n", (int)e.status, e.what()); return 0; }
            std::printf("probe unexpected success\n");
            return 1;
        }
        try { agent.train(env, 5, 0); std::printf("A no-throw\n"); } catch (const std::domain_error&) { std::printf("A eval_at==0 -> domain_error\n"); }
        auto [rewards, lengths, errors] = agent.train(env, n_episodes, n_episodes / 10);
        dump("A.rewards", rewards); dump("A.lengths", lengths); dump("A.errors", errors);
        auto [ev_rewards, ev_lengths] = agent.evaluate(env, 10);
        dump("A.eval_rewards", ev_rewards); dump("A.eval_lengths", ev_lengths);
        // the env on its own: terminated by evaluate() -> step() is Err(EnvNotReady); reset() revives it
        try { env.step({0, 0, 0}); std::printf("A no-throw\n"); } catch (const EnvNotReady&) { std::printf("A step after termination -> EnvNotReady\n"); }
        auto obs = env.reset();
        auto act = agent.get_action(obs);
        auto [obs2, rew2, term2] = env.step(act);
        std::printf("A.step %u %u %u | %u %u %u | %.1f %.1f %.1f\n", obs[0], obs[1], obs[2], obs2[0], obs2[1], obs2[2], rew2[0], rew2[1], rew2[2]);
        agent.reset();
        agent.set_action_selector(ucb);
        agent.set_future_q_value_func(expected_sarsa);
        auto [r2, l2, e2] = agent.train(env, 20, 2);
        dump("A2.lengths", l2); dump("A2.errors", e2);

        // --- BASELINE config C2's objects, f32
        FrozenLakeEnv lake(FrozenLakeEnv::MAP_8X8, true, 100);
        Batch b2;
        b2.n_agents = 40; b2.seed = 77; b2.real = RLB_REAL_F32;
        UniformEpsilonGreed eg2(1.0, Decay::sub(1.0 / (0.5 * 20.0)), 0.0);
        ElegibilityTracesAgent tracer(policy, 0.95, eg2, 0.5, sarsa, b2);
        auto [r3, l3, e3] = tracer.train(lake, 20, 2);
        std::printf("B.totals %" PRIu64 " %" PRIu64 " store=%u\n", tracer.last_train().train_steps, tracer.last_train().eval_steps,
                    rlb_engine_store_kind(tracer.engine()->get()));
        dump("B.lengths", l3);

        // --- bin/cliffwalking_model.rs in miniature: Dyna-Q around a second agent, two streams, f64
        CliffWalkingEnv cliff(100);
        Batch b3;
        b3.n_agents = 2; b3.seed = 0xD17A;
        UniformEpsilonGreed eg3(1.0, Decay::sub(1.0 / (0.5 * 30.0)), 0.0);
        OneStepAgent other(policy, 0.95, eg3, qlearning, b3);
        RandomModel model;
        {
            InternalModelAgent model_agent(other, model, 10);                              // bin/cliffwalking_model.rs:152-156
            auto [r4, l4, e4] = model_agent.train(cliff, 30, 3);
            dump("C.lengths", l4); dump("C.errors", e4);
            auto [len, ent] = model.entries();
            std::printf("C.model %u %u | %u %u %u %.1f\n", len[0], len[1], ent[0].obs, ent[0].action, ent[0].next_obs, ent[0].reward);
            auto info = model.get_info();
            std::printf("C.info %u %u %u %.1f\n", info.obs[0], info.action[0], info.next_obs[0], info.reward[0]);
        }
        auto [r5, l5, e5] = other.train(cliff, 5, 5);                                      // the borrow has ended: plain Q-learning again
        dump("C.after", l5);
        try { model.reset(); std::printf("C no-throw\n"); } catch (const std::logic_error&) { std::printf("C model unbound -> logic_error\n"); }

        // --- Agent::example / Env::render (agent.rs:143-163): one untrained episode per env, seed 0xE8A3, one agent each.
        // Every transcript line goes out as "D.<env>|<line with newlines as \n>".
        {
            Batch b4;
            b4.seed = 0xE8A3;
            UniformEpsilonGreed eg4(1.0, Decay::sub(1.0 / 15.0), 0.0);
            std::ostringstream sink;
            auto show = [&](const char* tag, Env& e) {
                OneStepAgent a(policy, 0.95, eg4, qlearning, b4);
                for (const std::string& line : a.example(e, sink)) {
                    std::string flat;
                    for (char ch : line) { if (ch == '\n') flat += "\\n"; else flat += ch; }
                    std::printf("D.%s|%s\n", tag, flat.c_str());
                }
            };
            TaxiEnv e1(100);
            FrozenLakeEnv e2_(FrozenLakeEnv::MAP_8X8, true, 100);
            CliffWalkingEnv e3_(100);
            BlackJackEnv e4_;
            show("taxi", e1); show("frozen_lake", e2_); show("cliff_walking", e3_); show("blackjack", e4_);
            try { other.example(cliff, sink); std::printf("D no-throw\n"); } catch (const std::logic_error&) { std::printf("D example with 2 agents -> logic_error\n"); }
        }
        // --- FrozenLakeEnv::new on the caller's own rows (frozen_lake.rs:48): two start cells, 5 x 7
        {
            FrozenLakeEnv lake2(std::vector<std::string>{"SFFFFFH", "FFHFFFF", "FFFFHFS", "HFFFFFF", "FFFHFFG"}, true, 40);
            Batch b6;
            b6.n_agents = 4; b6.seed = 0xF1A6;
            UniformEpsilonGreed eg6(1.0, Decay::sub(1.0 / (0.5 * 30.0)), 0.0);
            OneStepAgent walker(policy, 0.95, eg6, qlearning, b6);
            auto [r6, l6, e6] = walker.train(lake2, 30, 10);
            dump("F.lengths", l6); dump("F.rewards", r6);
        }
        // --- the reference's single trait object: ONE agent, training_error per STEP (agent.rs:98,117)
        {
            TaxiEnv taxi1(100);
            Batch b5;
            b5.n_agents = 1; b5.seed = 0x51;
            UniformEpsilonGreed eg5(1.0, Decay::sub(1.0 / (0.5 * 25.0)), 0.0);
            ElegibilityTracesAgent solo(policy, 0.95, eg5, 0.5, qlearning, b5);
            auto [r5, l5, e5] = solo.train(taxi1, 25, 5);
            dump("E.lengths", l5); dump("E.errors", e5);
        }
    } catch (const std::exception& e) {
        std::printf("FAILED: %s\n", e.what());
        return 1;
    }
    return 0;
}
