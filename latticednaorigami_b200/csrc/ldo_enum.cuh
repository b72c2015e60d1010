// ldo_enum.cuh — exact enumeration of small systems on the device.
//
// Reference: ConformationalEnumerator (src/enumerate.cpp:218-258, 378-664) — the recursive growth of every conformation
// of a scaffold with a given set of staples attached through a given set of growthpoints, each leaf weighted with
// exp(-E - bias) x multiplier. The staple sets and growthpoint sets (StapleEnumerator, GrowthpointEnumerator,
// enumerate.cpp:813-1133) are bookkeeping and stay on the host (ldo_sim.cpp); the host hands every growthpoint set to the
// device as an EnumJob.
//
// Parallelisation: the recursion is the reference's, executed warp-uniformly on the replica machinery (System::
// set_domain_config / unassign_domain on the warp's shared-memory state). The tree is cut at `split_depth` levels
// of enumerate_domain: a prefix is one (position, orientation) choice per level above the cut, 36^split_depth of them,
// dealt round-robin over the workers (warps); above the cut a worker follows only the branch its prefix names, below it
// enumerates everything. Levels that try a single orientation (bound, misbound, orientation-free placements) take
// orientation digit 0 only, and a leaf above the cut is counted by the prefix whose remaining digits are 0, so every
// leaf is visited exactly once. Every worker accumulates its own table of state weights; the host adds them up.
#pragma once

#include "ldo_moves.cuh"

namespace ldo {

#define LDO_ENUM_MAX_DOMAINS 32
#define LDO_ENUM_MAX_STAPLES 8
#define LDO_ENUM_MAX_KEYS 512
#define LDO_ENUM_MAX_OPS 6
#define LDO_ENUM_MAX_SPLIT 6
#define LDO_ENUM_MAX_IDENT 64

enum { ENUM_OVERCOUNT_MAX_TWO_DOMAIN = 0, ENUM_OVERCOUNT_MISBINDING_ONLY = 1 };

// One growthpoint set of one staple set. Chains are numbered 0 (scaffold), 1 + k (k-th staple of the job).
struct EnumJob {
    int n_staples;
    int staple_type[LDO_ENUM_MAX_STAPLES];
    // create_domains_stack (enumerate.cpp:591-637): domains in the order they are popped (index 0 first)
    int n_stack;
    short stack_chain[LDO_ENUM_MAX_DOMAINS], stack_d[LDO_ENUM_MAX_DOMAINS];
    // m_growthpoints: old domain -> new (staple) domain
    int n_gp;
    short gp_old_chain[LDO_ENUM_MAX_STAPLES], gp_old_d[LDO_ENUM_MAX_STAPLES];
    short gp_new_chain[LDO_ENUM_MAX_STAPLES], gp_new_d[LDO_ENUM_MAX_STAPLES];
    // m_identities_to_num_unassigned at the start of enumerate(), indexed by identity + LDO_ENUM_MAX_IDENT
    int ident_unassigned[2 * LDO_ENUM_MAX_IDENT + 1];
    int overcount;
    int n_out_ops;
    int out_op[LDO_ENUM_MAX_OPS];
    int split_depth;
    long long n_prefixes; // 36^split_depth
    int n_workers;
    // StapleConformationalEnumerator (enumerate.cpp:666-811): the scaffold keeps the configuration below, only the staples
    // attached to it are grown
    int staples_only;
    int scaf_pos[LDO_ENUM_MAX_DOMAINS][3];
    int scaf_ore[LDO_ENUM_MAX_DOMAINS]; // orientation codes
};

// Per worker: sums over its leaves (weights without the job's prefix factor, which the host applies)
struct EnumAcc {
    int n_keys;
    int status, status_detail;
    int pad;
    double z, avg_e, avg_b, n_configs;
    long long n_leaves;
    int keys[LDO_ENUM_MAX_KEYS][LDO_ENUM_MAX_OPS];
    double w[LDO_ENUM_MAX_KEYS];
};

template <class K>
struct Enumerator {
    Engine<K>& eng;
    const EnumJob& job;
    EnumAcc& acc;
    // per-lane working state (identical on every lane)
    short stack[LDO_ENUM_MAX_DOMAINS];
    int n_stack;
    short gp[K::D];
    short inv_gp[K::D]; // staples-only: staple domain -> the scaffold domain it grows off (m_inverse_growthpoints)
    V3 prev_ps[LDO_ENUM_MAX_DOMAINS];
    int n_prev;
    int id_un[2 * LDO_ENUM_MAX_IDENT + 1];
    double energy, mult;
    int dig_p[LDO_ENUM_MAX_SPLIT], dig_o[LDO_ENUM_MAX_SPLIT];

    LDO_HD Enumerator(Engine<K>& e, const EnumJob& j, EnumAcc& a): eng(e), job(j), acc(a) {}
    LDO_HD System<K>& sys() { return eng.sys; }
    LDO_HD int flat(int chain, int d) { return sys().dom_id(chain, d); }
    LDO_HD int& unassigned_of(int ident) { return id_un[ident + LDO_ENUM_MAX_IDENT]; }
    LDO_HD bool above_cut(int depth) const { return depth < job.split_depth; }
    LDO_HD int pop() { return stack[--n_stack]; }
    LDO_HD void push(int d) { stack[n_stack++] = (short)d; }
    LDO_HD bool ok() { return sys().S()->status == LDO_OK; }
    LDO_HD bool take_violation() {
        if (!sys().S()->constraints_violated) return false;
        sys().S()->constraints_violated = 0;
        return true;
    }

    // Builds the system of the job: scaffold and staples present, every domain unassigned (the state the
    // ConformationalEnumerator constructor and add_staple leave, enumerate.cpp:192-201, 260-289)
    LDO_HDN void build_system() {
        SysState<K>* s = sys().S();
        const SysConst* sc = &sys().SC();
        s->status = LDO_OK;
        s->status_detail = 0;
        s->constraints_violated = 0;
        s->weight_pass = 0;
        eng.BS()->pd_disabled = 1;
        sys().table_clear();
        for (int c = 0; c < K::C; c++) s->chain_used[c] = 0;
        for (int t = 0; t < K::T; t++) s->type_count[t] = 0;
        s->n_chains = 1;
        s->num_staples = 0;
        s->num_domains = 0;
        s->num_bound_pairs = 0;
        s->num_fully_bound_pairs = 0;
        s->num_self_bound_pairs = 0;
        s->num_stacked_pairs = 0;
        s->num_unassigned = 0;
        s->energy = 0;
        s->current_c_i = 0;
        int len = sc->type_len[0];
        s->chain_used[0] = 1;
        s->chain_uid[0] = 0;
        s->chain_type[0] = 0;
        s->chain_len[0] = (uint16_t)len;
        s->order[0] = 0;
        s->type_count[0] = 1;
        for (int i = 0; i < len; i++) {
            s->dom[i].k = 0;
            s->dom[i].ore = ORE_ZERO;
            s->dom[i].state = ST_UNASSIGNED;
            s->dom[i].link = chain_link_flags(i, len, sc->cyclic != 0);
            s->bound[i] = -1;
            s->ident[i] = sc->idents[sc->type_off[0] + i];
            s->dchain[i] = 0;
            s->dindex[i] = (uint16_t)i;
            s->num_domains++;
            s->num_unassigned++;
        }
        for (int k = 0; k < job.n_staples; k++) {
            int c = sys().add_chain(job.staple_type[k]);
            if (c != k + 1) sys().fail(LDO_ERR_INTERNAL, 700 + k);
        }
        if (job.staples_only) {
            // StapleConformationalEnumerator constructor: the scaffold domains are put back where the input has them
            for (int i = 0; i < len && ok(); i++) {
                sys().set_domain_config(i, v3(job.scaf_pos[i][0], job.scaf_pos[i][1], job.scaf_pos[i][2]), job.scaf_ore[i]);
            }
        }
        s->energy = 0; // the enumerator keeps its own sum (m_energy)
    }

    LDO_HDN void reset_job_state() {
        n_stack = 0;
        for (int i = job.n_stack - 1; i >= 0; i--) push(flat(job.stack_chain[i], job.stack_d[i])); // popped from the back
        for (int d = 0; d < K::D; d++) gp[d] = -1;
        for (int d = 0; d < K::D; d++) inv_gp[d] = -1;
        for (int g = 0; g < job.n_gp; g++) {
            int od = flat(job.gp_old_chain[g], job.gp_old_d[g]), nd = flat(job.gp_new_chain[g], job.gp_new_d[g]);
            gp[od] = (short)nd;
            if (job.staples_only && job.gp_old_chain[g] == 0) inv_gp[nd] = (short)od; // create_domains_stack, :799-807
        }
        for (int i = 0; i < 2 * LDO_ENUM_MAX_IDENT + 1; i++) id_un[i] = job.ident_unassigned[i];
        start_prefix();
    }
    // What a prefix starts from; the stack, the growthpoints and the identity counts are left as they were found by
    // every enumerate_prefix (the recursion is balanced), so only the running sums are set again
    LDO_HD void start_prefix() {
        n_prev = 0;
        // add_staple (enumerate.cpp:267-273): initiation energy and mean-field term of every staple of the set
        energy = 0;
        for (int k = 0; k < job.n_staples; k++) {
            if (sys().SC().apply_mean_field_cor) energy += log(6.0);
            energy += sys().TT().init_energy;
        }
        mult = 1;
    }

    // ---- overcount calculators (enumerate.cpp:83-159) ----
    LDO_HDN int count_involved_staples(int domain) {
        int involved = 0;
        bool on_scaffold = sys().chain(domain) == 0;
        if (!on_scaffold) involved++;
        int next = domain;
#pragma unroll 1
        while (!on_scaffold) {
            int f = sys().fwd(next);
            next = f < 0 ? sys().bac(next) : f;
            next = sys().bound(next);
            if (next < 0) {
                sys().fail(LDO_ERR_INTERNAL, 710); // the reference dereferences a null pointer here
                return involved;
            }
            if (sys().chain(next) == 0) on_scaffold = true;
            else involved++;
        }
        return involved;
    }
    LDO_HDN double overcount_multiplier(int domain, int other) {
        if (job.overcount == ENUM_OVERCOUNT_MAX_TWO_DOMAIN) {
            if (sys().chain(domain) == sys().chain(other)) return 1;
            int involved = count_involved_staples(domain) + count_involved_staples(other);
            return 1.0 / (involved + 1);
        }
        double m = 1;
        int next = sys().fwd(domain);
#pragma unroll 1
        while (next >= 0) {
            if (sys().state(next) == ST_BOUND) m += 1;
            else if (sys().state(next) == ST_UNASSIGNED) return 1;
            next = sys().fwd(next);
        }
        next = sys().bac(domain);
#pragma unroll 1
        while (next >= 0) {
            if (sys().state(next) == ST_BOUND) m += 1;
            else if (sys().state(next) == ST_UNASSIGNED) return 1;
            next = sys().bac(next);
        }
        return m;
    }

    // ---- calc_and_save_weights (enumerate.cpp:639-664) ----
    LDO_HDN void save_weights() {
        eng.update_move_params();
        eng.calc_move_bias();
        double conf_bias = eng.total_bias();
        double weight = exp(-energy - conf_bias) * mult;
        if (LDO_LANE == 0) {
            acc.n_configs += mult;
            acc.n_leaves += 1;
            acc.avg_e += energy * weight;
            acc.avg_b += conf_bias * weight;
            acc.z += weight;
            int n = acc.n_keys, hit = -1;
#pragma unroll 1
            for (int k = 0; k < n && hit < 0; k++) {
                bool same = true;
                for (int i = 0; i < job.n_out_ops; i++) same = same && acc.keys[k][i] == eng.BS()->op_val[job.out_op[i]];
                if (same) hit = k;
            }
            if (hit < 0) {
                if (n >= LDO_ENUM_MAX_KEYS) {
                    acc.status = LDO_ERR_CAPACITY;
                    acc.status_detail = 720;
                }
                else {
                    for (int i = 0; i < job.n_out_ops; i++) acc.keys[n][i] = eng.BS()->op_val[job.out_op[i]];
                    acc.w[n] = weight;
                    acc.n_keys = n + 1;
                }
            }
            else {
                acc.w[hit] += weight;
            }
        }
        LDO_SYNCWARP();
    }

    // ---- the recursion (enumerate.cpp:378-589) ----
    LDO_HDN void enumerate_domain(int domain, V3 p_prev, int depth) {
#pragma unroll 1
        for (int pi = 0; pi < 6; pi++) {
            if (!ok()) break;
            if (above_cut(depth) && pi != dig_p[depth]) continue;
            V3 p_new = p_prev + ore_vec(pi);
            bool is_growthpoint = gp[domain] >= 0;
            int j = sys().occupant(p_new);
            int occ = j < 0 ? ST_UNASSIGNED : sys().state(j);
            bool is_occupied = occ == ST_UNBOUND;
            bool is_bound = occ == ST_BOUND || occ == ST_MISBOUND;
            if ((is_growthpoint && is_occupied) || is_bound) continue;
            else if (is_growthpoint) set_growthpoint_domains(domain, p_new, depth);
            else if (is_occupied) set_bound_domain(domain, p_new, j, depth);
            else set_unbound_domain(domain, p_new, depth);
        }
        push(domain);
    }
    LDO_HDN void set_growthpoint_domains(int domain, V3 p_new, int depth) {
        bool terminal = false;
        int f = sys().fwd(domain), b = sys().bac(domain);
        if (f >= 0 && sys().state(f) == ST_UNASSIGNED) {
            if (b >= 0) terminal = true;
        }
        else {
            if (f >= 0) terminal = true;
        }
        if (terminal) prev_ps[n_prev++] = p_new;
        int bound_domain = pop();
        if (sys().ident(domain) == -sys().ident(bound_domain)) set_comp_growthpoint_domains(domain, bound_domain, p_new, depth);
        else set_mis_growthpoint_domains(domain, bound_domain, p_new, depth);
        push(bound_domain);
        if (terminal) n_prev--;
    }
    LDO_HDN void set_comp_growthpoint_domains(int domain, int bound_domain, V3 p_new, int depth) {
#pragma unroll 1
        for (int oi = 0; oi < 6; oi++) {
            if (!ok()) break;
            if (above_cut(depth) && oi != dig_o[depth]) continue;
            energy += sys().set_domain_config(domain, p_new, oi);
            energy += sys().set_domain_config(bound_domain, p_new, oi ^ 1);
            if (take_violation()) {
                energy += sys().unassign_domain(domain);
                continue;
            }
            unassigned_of(sys().ident(domain)) -= 1;
            unassigned_of(sys().ident(bound_domain)) -= 1;
            grow_next_domain(bound_domain, p_new, depth);
            unassigned_of(sys().ident(domain)) += 1;
            unassigned_of(sys().ident(bound_domain)) += 1;
            energy += sys().unassign_domain(domain);
        }
    }
    LDO_HDN void set_mis_growthpoint_domains(int domain, int bound_domain, V3 p_new, int depth) {
        if (above_cut(depth) && dig_o[depth] != 0) return;
        energy += sys().set_domain_config(domain, p_new, 0);
        energy += sys().set_domain_config(bound_domain, p_new, 1);
        mult *= 6;
        unassigned_of(sys().ident(domain)) -= 1;
        unassigned_of(sys().ident(bound_domain)) -= 1;
        grow_next_domain(bound_domain, p_new, depth);
        unassigned_of(sys().ident(domain)) += 1;
        unassigned_of(sys().ident(bound_domain)) += 1;
        energy += sys().unassign_domain(domain);
        mult /= 6;
    }
    LDO_HDN void set_bound_domain(int domain, V3 p_new, int occ_domain, int depth) {
        if (above_cut(depth) && dig_o[depth] != 0) return;
        int oc = sys().orc(occ_domain);
        int o_new = oc < ORE_ZERO ? (oc ^ 1) : oc;
        energy += sys().set_domain_config(domain, p_new, o_new);
        if (take_violation()) return;
        double pos_multiplier = overcount_multiplier(domain, occ_domain);
        mult *= pos_multiplier;
        unassigned_of(sys().ident(domain)) -= 1;
        grow_next_domain(domain, p_new, depth);
        unassigned_of(sys().ident(domain)) += 1;
        mult /= pos_multiplier;
    }
    LDO_HDN void set_unbound_domain(int domain, V3 p_new, int depth) {
        if (job.staples_only) {
            // StapleConformationalEnumerator::set_unbound_domain (:781-791): one orientation-free placement
            if (above_cut(depth) && dig_o[depth] != 0) return;
            energy += sys().set_domain_config(domain, p_new, ORE_ZERO);
            mult *= 6;
            grow_next_domain(domain, p_new, depth);
            mult /= 6;
            return;
        }
        // all orientations only when a twist constraint is possible (enumerate.cpp:525-549)
        if (unassigned_of(-sys().ident(domain)) == 0) {
            if (above_cut(depth) && dig_o[depth] != 0) return;
            unassigned_of(sys().ident(domain)) -= 1;
            energy += sys().set_domain_config(domain, p_new, ORE_ZERO);
            mult *= 6;
            grow_next_domain(domain, p_new, depth);
            unassigned_of(sys().ident(domain)) += 1;
            mult /= 6;
        }
        else {
#pragma unroll 1
            for (int oi = 0; oi < 6; oi++) {
                if (!ok()) break;
                if (above_cut(depth) && oi != dig_o[depth]) continue;
                energy += sys().set_domain_config(domain, p_new, oi);
                if (take_violation()) continue;
                unassigned_of(sys().ident(domain)) -= 1;
                grow_next_domain(domain, p_new, depth);
                unassigned_of(sys().ident(domain)) += 1;
            }
        }
    }
    // a leaf above the cut belongs to the prefix whose remaining digits are zero
    LDO_HD bool leaf_is_mine(int depth) const {
        bool mine = true;
        for (int l = depth + 1; l < job.split_depth; l++) mine = mine && dig_p[l] == 0 && dig_o[l] == 0;
        return mine;
    }
    // StapleConformationalEnumerator::grow_off_scaffold (:754-779): the staple domain bound to its scaffold domain
    LDO_HDN void grow_off_scaffold(int next_domain, int depth) {
        int sd = inv_gp[next_domain];
        V3 p_new = sys().pos(sd);
        int oc = sys().orc(sd);
        energy += sys().set_domain_config(next_domain, p_new, oc < ORE_ZERO ? (oc ^ 1) : oc);
        if (take_violation()) {
            push(next_domain);
            return;
        }
        if (n_stack > 0) {
            int nn = pop();
            if (inv_gp[nn] >= 0) grow_off_scaffold(nn, depth);
            else enumerate_domain(nn, p_new, depth + 1);
            push(next_domain);
        }
        // (a set that ends here is not counted: the reference saves no weight on this path)
        energy += sys().unassign_domain(next_domain);
    }
    LDO_HDN void grow_next_domain(int domain, V3 p_new, int depth) {
        if (job.staples_only) {
            // StapleConformationalEnumerator::grow_next_domain (:729-752)
            if (n_stack > 0) {
                int next = pop();
                if (inv_gp[next] >= 0) grow_off_scaffold(next, depth);
                else enumerate_domain(next, p_new, depth + 1);
            }
            else if (leaf_is_mine(depth)) {
                save_weights();
            }
            energy += sys().unassign_domain(domain);
            return;
        }
        if (n_stack > 0) {
            int next = pop();
            V3 new_p_prev;
            bool terminal = sys().chain(domain) != sys().chain(next);
            if (terminal) new_p_prev = prev_ps[--n_prev];
            else new_p_prev = p_new;
            enumerate_domain(next, new_p_prev, depth + 1);
            if (terminal) prev_ps[n_prev++] = new_p_prev;
        }
        else {
            if (leaf_is_mine(depth)) {
                if (sys().SC().cyclic) {
                    int last = sys().S()->chain_len[0] - 1;
                    if (abssum(sys().pos(0) - sys().pos(last)) == 1) save_weights();
                }
                else {
                    save_weights();
                }
            }
        }
        energy += sys().unassign_domain(domain);
    }

    // ConformationalEnumerator::enumerate (enumerate.cpp:218-258) for one prefix
    LDO_HDN void enumerate_prefix(long long prefix) {
        for (int l = 0; l < LDO_ENUM_MAX_SPLIT; l++) {
            dig_p[l] = 0;
            dig_o[l] = 0;
        }
        for (int l = job.split_depth - 1; l >= 0; l--) {
            int digit = (int)(prefix % 36);
            prefix /= 36;
            dig_p[l] = digit / 6;
            dig_o[l] = digit % 6;
        }
        start_prefix();
        if (job.staples_only) {
            enumerate_staples_prefix();
            return;
        }
        int starting = pop();
        V3 p_new = v3(0, 0, 0);
        unassigned_of(sys().ident(starting)) -= 1;
        sys().set_domain_config(starting, p_new, 0);
        bool is_growthpoint = gp[starting] >= 0;
        int next = pop();
        if (is_growthpoint) {
            prev_ps[n_prev++] = p_new;
            unassigned_of(sys().ident(next)) -= 1;
            energy += sys().set_domain_config(next, p_new, 1);
            int next_next = pop();
            enumerate_domain(next_next, p_new, 0);
            unassigned_of(sys().ident(next)) += 1;
            energy += sys().unassign_domain(next);
            n_prev--;
            push(next); // (the reference rebuilds its stack for every call instead)
        }
        else {
            enumerate_domain(next, p_new, 0);
        }
        unassigned_of(sys().ident(starting)) += 1;
        sys().unassign_domain(starting);
        push(starting);
    }

    // StapleConformationalEnumerator::enumerate (:687-717)
    LDO_HDN void enumerate_staples_prefix() {
        int n0 = n_stack;
        if (n_stack == 0) {
            if (leaf_is_mine(-1)) save_weights();
            return;
        }
        int starting = pop();
        int sd = inv_gp[starting];
        V3 p_new = sys().pos(sd);
        int oc = sys().orc(sd);
        energy += sys().set_domain_config(starting, p_new, oc < ORE_ZERO ? (oc ^ 1) : oc);
        if (n_stack > 0) {
            int next = pop();
            if (inv_gp[next] >= 0) grow_off_scaffold(next, -1);
            else enumerate_domain(next, p_new, 0);
        }
        else if (leaf_is_mine(-1)) {
            save_weights();
        }
        energy += sys().unassign_domain(starting);
        // the reference rebuilds its stack for every call; here it is put back as it was
        n_stack = n0;
    }

    LDO_HDN void run(int worker) {
        build_system();
        reset_job_state();
#pragma unroll 1
        for (long long q = worker; q < job.n_prefixes && ok(); q += job.n_workers) enumerate_prefix(q);
        if (LDO_LANE == 0 && !ok()) {
            acc.status = sys().S()->status;
            acc.status_detail = sys().S()->status_detail;
        }
    }
};

} // namespace ldo
