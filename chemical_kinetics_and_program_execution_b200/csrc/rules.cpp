// Host-side rule front end (see rules.h).
#include "rules.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <stdexcept>

namespace tapes {

namespace {

const int kMaxSide = 40;  // cells a view may extend to either side of cell 0

// A contiguous symbolic view [-left, right-1] holding both the original and the adjusted symbol
// of every uncovered cell (tm.scm:41-55 t-view, tm.scm:207-216 tv-pair).
struct Window {
  int left = 0, right = 0;
  uint16_t orig[2 * kMaxSide];
  uint16_t adj[2 * kMaxSide];
  bool covers(int cell) const { return cell >= -left && cell < right; }
  int width() const { return left + right; }
  uint64_t digits(const uint16_t* a, int first, int count, int A) const {
    uint64_t v = 0;
    for (int i = 0; i < count; ++i) v = v * (uint64_t)A + a[kMaxSide + first + i];
    return v;
  }
};

struct Fork { int ways; };  // thrown when the decision list is exhausted at a split point

// Runs a body against a fixed list of split decisions, recording the probability steps.
class Replayer : public Machine {
 public:
  Replayer(int A, int k, const std::vector<int>& decisions) : A_(A), k_(k), decisions_(decisions) {}

  int read(Tape t, int cell) override {
    Window& w = win_[t];
    while (!w.covers(cell)) {
      if (cell >= kMaxSide || cell < -kMaxSide) throw std::runtime_error("tape view too long");
      int s = next_decision(A_);
      // One-cell unfolding toward `cell` (tm.scm:482-588): L = min(k, visible + 1), the context
      // is the L-1 cells adjacent to the new cell, always taken from the ORIGINAL view (490).
      const bool to_right = cell >= 0;
      const int L = std::min(k_, w.width() + 1);
      uint64_t context = to_right ? w.digits(w.orig, w.right - (L - 1), L - 1, A_)
                                  : w.digits(w.orig, -w.left, L - 1, A_);
      uint64_t pow_lm1 = 1;
      for (int i = 0; i < L - 1; ++i) pow_lm1 *= (uint64_t)A_;
      uint64_t extended = to_right ? context * (uint64_t)A_ + (uint64_t)s
                                   : (uint64_t)s * pow_lm1 + context;
      Step st;
      st.kind = Step::UNFOLD;
      st.length = (uint8_t)L;
      st.long_index = (uint32_t)extended;
      st.short_index = (uint32_t)context;
      st.prob = 0.0;
      steps_.push_back(st);
      int pos = to_right ? w.right++ : -(++w.left);
      w.orig[kMaxSide + pos] = (uint16_t)s;
      w.adj[kMaxSide + pos] = (uint16_t)s;
    }
    return w.adj[kMaxSide + cell];  // programs observe their own writes (tm.scm:759)
  }

  void write(Tape t, int cell, int symbol) override {
    (void)read(t, cell);  // a write first forces the cell to be unfolded (tm.scm:787-792)
    win_[t].adj[kMaxSide + cell] = (uint16_t)symbol;
  }

  int pick(const double* weights, int n) override {
    double total = 0.0;
    for (int i = 0; i < n; ++i) total = total + weights[i];
    int j = next_decision(n);
    Step st;
    st.kind = Step::CHOICE;
    st.length = 0;
    st.long_index = st.short_index = 0;
    st.prob = weights[j] / total;
    steps_.push_back(st);
    return j;
  }

  const std::vector<Step>& steps() const { return steps_; }
  const Window& window(int t) const { return win_[t]; }

 private:
  int next_decision(int ways) {
    if (pos_ >= decisions_.size()) throw Fork{ways};
    return decisions_[pos_++];
  }
  int A_, k_;
  const std::vector<int>& decisions_;
  size_t pos_ = 0;
  Window win_[2];
  std::vector<Step> steps_;
};

std::map<std::string, Problem>& registry() {
  static std::map<std::string, Problem> r;
  return r;
}

}  // namespace

RuleTable enumerate_rules(const Problem& problem, int cl_k) {
  RuleTable table;
  table.alphabet = problem.alphabet;
  table.cl_k = cl_k;
  double states = 1;
  for (int i = 0; i < cl_k; ++i) states *= problem.alphabet;
  if (cl_k < 1 || states >= 4294967296.0)
    throw std::runtime_error("A^cl_k must be below 2^32");

  std::vector<std::vector<int>> todo;
  todo.push_back({});
  while (!todo.empty()) {
    std::vector<int> decisions = std::move(todo.back());
    todo.pop_back();
    table.worlds_walked++;
    Replayer m(problem.alphabet, cl_k, decisions);
    try {
      problem.body(m);
    } catch (const Fork& f) {
      for (int c = f.ways - 1; c >= 0; --c) {  // pushed high-to-low so that option 0 is walked first
        std::vector<int> d = decisions;
        d.push_back(c);
        todo.push_back(std::move(d));
      }
      continue;
    }
    table.leaf_worlds++;
    FluxRule rule;
    bool any = false;
    for (int t = 0; t < 2; ++t) {
      const Window& w = m.window(t);
      Seed& sd = rule.tape[t];
      sd.length = w.width();
      double bits = 0;
      for (int i = 0; i < sd.length; ++i) bits += std::log2((double)problem.alphabet);
      if (bits > 63.0) throw std::runtime_error("tape view index exceeds 64 bits");
      sd.orig = w.digits(w.orig, -w.left, sd.length, problem.alphabet);
      sd.adjusted = w.digits(w.adj, -w.left, sd.length, problem.alphabet);
      any = any || sd.changed();
    }
    if (!any) continue;
    rule.steps = m.steps();
    table.rules.push_back(std::move(rule));
  }
  return table;
}

std::vector<double> flux_rule_costs(const RuleTable& table) {
  const int A = table.alphabet, k = table.cl_k;
  std::vector<double> costs;
  for (const FluxRule& r : table.rules) {
    double cost = 0.0;
    for (int t = 0; t < 2; ++t) {
      const Seed& sd = r.tape[t];
      if (!sd.changed()) continue;
      // changed[pos]: cell pos of the view (0 = left-most) differs between the two views
      std::vector<char> changed((size_t)sd.length, 0);
      int first = sd.length, last = -1;
      uint64_t o = sd.orig, a = sd.adjusted;
      for (int pos = sd.length - 1; pos >= 0; --pos) {
        if (o % (uint64_t)A != a % (uint64_t)A) { changed[(size_t)pos] = 1; first = std::min(first, pos); last = std::max(last, pos); }
        o /= (uint64_t)A; a /= (uint64_t)A;
      }
      for (int start = first - k + 1; start <= last; ++start) {
        bool hit = false;
        for (int pos = std::max(start, 0); pos < std::min(start + k, sd.length); ++pos) hit = hit || changed[(size_t)pos];
        if (!hit) continue;
        const int inside = std::max(0, std::min(start + k, sd.length) - std::max(start, 0));
        cost += std::pow((double)A, k - inside);
      }
    }
    costs.push_back(cost);
  }
  return costs;
}

std::vector<int> deal_flux_rules(const std::vector<double>& costs, int n_parts) {
  if (n_parts < 1) throw std::runtime_error("need at least one part");
  std::vector<size_t> order(costs.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return costs[x] > costs[y]; });
  std::vector<double> load((size_t)n_parts, 0.0);
  std::vector<int> owner(costs.size(), 0);
  for (size_t r : order) {
    int best = 0;
    for (int g = 1; g < n_parts; ++g)
      if (load[(size_t)g] < load[(size_t)best]) best = g;
    owner[r] = best;
    load[(size_t)best] += costs[r];
  }
  return owner;
}

RuleTable rule_table_part(const RuleTable& table, int part, int n_parts) {
  if (part < 0 || part >= n_parts) throw std::runtime_error("part outside 0..n_parts-1");
  const std::vector<int> owner = deal_flux_rules(flux_rule_costs(table), n_parts);
  RuleTable out;
  out.alphabet = table.alphabet;
  out.cl_k = table.cl_k;
  out.worlds_walked = table.worlds_walked;
  out.leaf_worlds = table.leaf_worlds;
  for (size_t r = 0; r < table.rules.size(); ++r)
    if (owner[r] == part) out.rules.push_back(table.rules[r]);
  return out;
}

const Problem* find_problem(const std::string& tag) {
  auto it = registry().find(tag);
  return it == registry().end() ? nullptr : &it->second;
}

void register_problem(const std::string& tag, int alphabet, Body body) {
  Problem p;
  p.alphabet = alphabet;
  p.body = std::move(body);
  registry()[tag] = std::move(p);
}

std::vector<std::string> registered_tags() {
  std::vector<std::string> tags;
  for (auto& kv : registry()) tags.push_back(kv.first);
  return tags;
}

Body body_from_rewrite_rules(std::vector<RewriteRule> rules) {
  for (const RewriteRule& r : rules)
    if (r.span < 1 || r.span > 4) throw std::runtime_error("rewrite rule span must be 1..4");
  return [rules](Machine& m) {
    std::vector<double> weights(rules.size());
    for (size_t i = 0; i < rules.size(); ++i) weights[i] = rules[i].select_weight;
    const RewriteRule& r = rules[(size_t)m.pick(weights.data(), (int)weights.size())];
    const Tape own = r.tape ? DATA_TAPE : PROGRAM_TAPE;
    const Tape other = r.tape ? PROGRAM_TAPE : DATA_TAPE;
    if (r.catalyst >= 0 && m.read(other, 0) != r.catalyst) return;
    for (int c = 0; c < r.span; ++c)
      if (m.read(own, c) != r.pattern[c]) return;
    const double accept[2] = {r.rate, 1 - r.rate};
    if (m.pick(accept, 2) != 0) return;
    for (int c = 0; c < r.span; ++c)
      if (r.replacement[c] != r.pattern[c]) m.write(own, c, r.replacement[c]);
  };
}

Body body_from_program(ProgramTree t, int alphabet) {
  const size_t n = t.kind.size();
  if (n == 0) throw std::runtime_error("program tree has no nodes");
  if (t.a.size() != n || t.b.size() != n || t.c.size() != n || t.first_child.size() != n || t.first_weight.size() != n)
    throw std::runtime_error("program tree arrays differ in length");
  for (size_t i = 0; i < n; ++i) {
    int fan = 0;
    switch (t.kind[i]) {
      case ProgramTree::END: break;
      case ProgramTree::READ: fan = alphabet; break;
      case ProgramTree::WRITE:
        fan = 1;
        if (t.c[i] < 0 || t.c[i] >= alphabet) throw std::runtime_error("program writes a symbol outside the alphabet");
        break;
      case ProgramTree::PICK:
        fan = t.a[i];
        if (fan < 1) throw std::runtime_error("choice without options");
        if (t.first_weight[i] < 0 || (size_t)t.first_weight[i] + (size_t)fan > t.weight.size())
          throw std::runtime_error("choice weights outside the weight array");
        break;
      default: throw std::runtime_error("unknown program node kind");
    }
    if (t.kind[i] == ProgramTree::READ || t.kind[i] == ProgramTree::WRITE) {
      if (t.a[i] != 0 && t.a[i] != 1) throw std::runtime_error("tape must be 0 (program) or 1 (data)");
      if (t.b[i] <= -kMaxSide || t.b[i] >= kMaxSide) throw std::runtime_error("cell index too far from the head");
    }
    if (fan) {
      if (t.first_child[i] < 0 || (size_t)t.first_child[i] + (size_t)fan > t.child.size())
        throw std::runtime_error("children outside the child array");
      for (int j = 0; j < fan; ++j) {
        const int32_t ch = t.child[(size_t)t.first_child[i] + (size_t)j];
        if (ch <= (int32_t)i || (size_t)ch >= n) throw std::runtime_error("a child must come after its parent");
      }
    }
  }
  return [t](Machine& m) {
    size_t i = 0;
    for (;;) {
      const size_t at = (size_t)t.first_child[i];
      switch (t.kind[i]) {
        case ProgramTree::END: return;
        case ProgramTree::READ: i = (size_t)t.child[at + (size_t)m.read(t.a[i] ? DATA_TAPE : PROGRAM_TAPE, t.b[i])]; break;
        case ProgramTree::WRITE: m.write(t.a[i] ? DATA_TAPE : PROGRAM_TAPE, t.b[i], t.c[i]); i = (size_t)t.child[at]; break;
        default: i = (size_t)t.child[at + (size_t)m.pick(&t.weight[(size_t)t.first_weight[i]], t.a[i])]; break;
      }
    }
  };
}

namespace {

struct TraceFork { int ways; };

struct TraceOp {
  int32_t kind, a, b, c;
  std::vector<double> weights;
  bool operator==(const TraceOp& o) const {
    return kind == o.kind && a == o.a && b == o.b && c == o.c && weights == o.weights;
  }
};

// Records what a body does at the level of its own calls (not of the cell-by-cell unfolding).
class Tracer : public Machine {
 public:
  Tracer(int A, const std::vector<int>& decisions) : A_(A), decisions_(decisions) {}
  int read(Tape t, int cell) override {
    auto it = known_[t].find(cell);
    if (it != known_[t].end()) return it->second;
    ops.push_back(TraceOp{ProgramTree::READ, (int32_t)t, cell, 0, {}});
    const int s = decide(A_);
    known_[t][cell] = s;
    return s;
  }
  void write(Tape t, int cell, int symbol) override {
    ops.push_back(TraceOp{ProgramTree::WRITE, (int32_t)t, cell, symbol, {}});
    known_[t][cell] = symbol;
  }
  int pick(const double* weights, int n) override {
    ops.push_back(TraceOp{ProgramTree::PICK, n, 0, 0, std::vector<double>(weights, weights + n)});
    return decide(n);
  }
  std::vector<TraceOp> ops;

 private:
  int decide(int ways) {
    if (pos_ >= decisions_.size()) throw TraceFork{ways};
    return decisions_[pos_++];
  }
  int A_;
  const std::vector<int>& decisions_;
  size_t pos_ = 0;
  std::map<int, int> known_[2];
};

struct TrieNode {
  bool has_op = false;
  TraceOp op;
  std::vector<int> children;  // indices into the trie
};

}  // namespace

ProgramTree trace_body(const Body& body, int alphabet, size_t max_nodes) {
  std::vector<TrieNode> trie(1);
  std::vector<std::vector<int>> todo;
  todo.push_back({});
  while (!todo.empty()) {
    std::vector<int> decisions = std::move(todo.back());
    todo.pop_back();
    Tracer m(alphabet, decisions);
    try {
      body(m);
    } catch (const TraceFork& f) {
      for (int c = f.ways - 1; c >= 0; --c) {
        std::vector<int> d = decisions;
        d.push_back(c);
        todo.push_back(std::move(d));
      }
      continue;
    }
    size_t node = 0, used = 0;
    for (const TraceOp& op : m.ops) {
      if (!trie[node].has_op) {
        const int fan = op.kind == ProgramTree::READ ? alphabet : (op.kind == ProgramTree::WRITE ? 1 : op.a);
        trie[node].has_op = true;
        trie[node].op = op;
        for (int j = 0; j < fan; ++j) {
          trie[node].children.push_back((int)trie.size());
          trie.emplace_back();
        }
        if (trie.size() > max_nodes) throw std::runtime_error("program tree too large");
      } else if (!(trie[node].op == op)) {
        throw std::runtime_error("the body is not deterministic");
      }
      node = (size_t)trie[node].children[op.kind == ProgramTree::WRITE ? 0 : (size_t)decisions[used++]];
    }
    if (trie[node].has_op) throw std::runtime_error("the body is not deterministic");
  }
  // flatten in preorder; every run ends in the shared END node, which comes last
  std::vector<int> id(trie.size(), -1), order, stack(1, 0);
  while (!stack.empty()) {
    const int n = stack.back();
    stack.pop_back();
    if (!trie[(size_t)n].has_op) continue;
    id[(size_t)n] = (int)order.size();
    order.push_back(n);
    for (size_t j = trie[(size_t)n].children.size(); j-- > 0;) stack.push_back(trie[(size_t)n].children[j]);
  }
  const int end_id = (int)order.size();
  ProgramTree t;
  const size_t n = order.size() + 1;
  t.kind.assign(n, ProgramTree::END); t.a.assign(n, 0); t.b.assign(n, 0); t.c.assign(n, 0);
  t.first_child.assign(n, 0); t.first_weight.assign(n, 0);
  for (size_t i = 0; i < order.size(); ++i) {
    const TrieNode& nd = trie[(size_t)order[i]];
    t.kind[i] = nd.op.kind; t.a[i] = nd.op.a; t.b[i] = nd.op.b; t.c[i] = nd.op.c;
    t.first_child[i] = (int32_t)t.child.size();
    for (int ch : nd.children) t.child.push_back(trie[(size_t)ch].has_op ? id[(size_t)ch] : end_id);
    if (nd.op.kind == ProgramTree::PICK) {
      t.first_weight[i] = (int32_t)t.weight.size();
      t.weight.insert(t.weight.end(), nd.op.weights.begin(), nd.op.weights.end());
    }
  }
  return t;
}

}  // namespace tapes
