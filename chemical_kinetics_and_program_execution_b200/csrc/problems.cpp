// Built-in problems, registered under the reference's tag strings
// (framework/problems.scm:18-639; canary: framework/tapes_py_interface.scm:105-109).
// Each body states the same decision procedure as the Scheme body it cites; symbols are alphabet
// indices in the order of the reference's alphabet vectors.
#include <cmath>
#include <stdexcept>

#include "rules.h"

namespace tapes {

namespace {

const Tape P = PROGRAM_TAPE;
const Tape D = DATA_TAPE;

bool coin(Machine& m) {  // (choose '((1.0 #t) (1.0 #f)))
  const double even[2] = {1.0, 1.0};
  return m.pick(even, 2) == 0;
}

// ---- ex1 / canary: alphabet (A B); a B under the data head decays to A.  prob.scm:22-26 ----
void decay(Machine& m) {
  enum { A_ = 0, B_ = 1 };
  if (m.read(D, 0) == B_) m.write(D, 0, A_);
}

// ---- ex2: alphabet (D U).  prob.scm:30-55 ----
void ferromagnet(Machine& m) {
  enum { DOWN = 0, UP = 1 };
  const double J = 1.0, h = -0.25, beta = 1.0;
  const int mid = m.read(D, 0);
  const int lft = m.read(D, -1);
  const int rgt = m.read(D, +1);
  const int bonds = (lft == mid ? 1 : -1) + (mid == rgt ? 1 : -1);
  const double coupling = std::exp(-(beta * J * (4 + 2 * bonds)));
  // the field suppresses flips away from the direction it favours (prob.scm:43-50)
  const bool aligned = (h > 0) == (mid == UP);
  const double field = aligned ? std::exp(-(2 * beta * std::fabs(h))) : 1.0;
  const double flip = coupling * field;
  const double options[2] = {flip, 1 - flip};
  if (m.pick(options, 2) == 0) m.write(D, 0, mid == UP ? DOWN : UP);
}

// ---- ex3 family: alphabet (O A M N).  prob.scm:59-181 ----
namespace copoly {
enum { O = 0, A = 1, M = 2, N = 3 };

bool isolated_on_program_tape(Machine& m) { return m.read(P, -1) == O && m.read(P, +1) == O; }

bool pairs_up(int a, int b) {
  auto amine = [](int s) { return s == M || s == N; };
  return (a == A && amine(b)) || (b == A && amine(a));
}

// The condensation step shared by all three variants (prob.scm:63-85, 93-126, 160-181).
void condense(Machine& m, int monomer, bool alternation_bias) {
  if (!isolated_on_program_tape(m)) return;
  const int under = m.read(D, 0);
  if (!pairs_up(monomer, under)) return;
  const int side = coin(m) ? -1 : +1;  // '((1.0 -1) (1.0 +1))
  if (m.read(D, side) != O) return;
  if (m.read(D, 2 * side) != O) return;
  if (alternation_bias && monomer != A && m.read(D, -side) == monomer) {
    const double reject[2] = {75.0, 25.0};  // prob.scm:120
    if (m.pick(reject, 2) == 0) return;
  }
  m.write(P, 0, O);
  m.write(D, side, monomer);
}

void plain(Machine& m) {
  const int p0 = m.read(P, 0);
  if (p0 != O) condense(m, p0, false);
}
void alternating(Machine& m) {
  const int p0 = m.read(P, 0);
  if (p0 != O) condense(m, p0, true);
}
void reversible(Machine& m) {  // prob.scm:130-181
  const int p0 = m.read(P, 0);
  if (p0 != O) { condense(m, p0, false); return; }
  if (!isolated_on_program_tape(m)) return;
  const int unit = m.read(D, 0);
  if (unit == O) return;
  const int right = m.read(D, 1);
  const int left = m.read(D, -1);
  const int neighbours = (left != O) + (right != O);
  if (neighbours != 1) return;  // only chain ends come off
  const double release[2] = {1.0, 50.0};
  if (m.pick(release, 2) != 0) return;
  m.write(P, 0, unit);
  m.write(D, 0, O);
}
}  // namespace copoly

// ---- ex4 family: alphabet (A B C D I O P X S [E]).  prob.scm:186-434 ----
namespace turing {
enum { cA = 0, cB = 1, cC = 2, cD = 3, bI = 4, bO = 5, mP = 6, mX = 7, mS = 8, mE = 9 };

bool bit(int s) { return s == bI || s == bO; }
int random_bit(Machine& m) { return coin(m) ? bI : bO; }  // '((1.0 I) (1.0 O))

// cursor A writes I and becomes B, B writes O and becomes C, C writes I and becomes D
struct Advance { int from, leaves, becomes; };
const Advance kAdvance[3] = {{cA, bI, cB}, {cB, bO, cC}, {cC, bI, cD}};

// prob.scm:194-245 and 255-305; the variants differ only in the reverse-suppression choice.
void machine(Machine& m, const double* suppress_weights, int accept_option) {
  const int fuel = m.read(P, 0);
  if (fuel == mP && coin(m)) {
    const int cursor = m.read(D, 0);
    for (const Advance& a : kAdvance) {
      if (cursor == a.from && bit(m.read(D, 1)) && bit(m.read(D, 2))) {
        m.write(P, 0, mX);
        m.write(D, 0, a.leaves);
        m.write(D, 1, a.becomes);
        return;
      }
    }
    return;
  }
  if (fuel != mX) return;
  const int cursor = m.read(D, 0);
  if (cursor != cB && cursor != cC && cursor != cD) return;
  if (!bit(m.read(D, -1)) || !bit(m.read(D, -2))) return;
  const int expected = cursor == cC ? bO : bI;  // what the forward step would have left behind
  if (m.read(D, -1) != expected) return;
  if (m.pick(suppress_weights, 2) != accept_option) return;
  m.write(P, 0, mP);
  m.write(D, 0, random_bit(m));
  m.write(D, -1, cursor - 1);  // B->A, C->B, D->C
}
void base(Machine& m) {  // prob.scm:187-190: ((0.95 #f) (0.05 #t))
  const double f = 0.05;
  const double w[2] = {1.0 - f, f};
  machine(m, w, 1);
}
void unsuppressed(Machine& m) {  // prob.scm:248-251: ((1.0 #t) (0.0 #f))
  const double f = 0.0;
  const double w[2] = {1.0 - f, f};
  machine(m, w, 0);
}

// prob.scm:307-434.
struct Enthalpies {
  double fwd[3][2], rev[3][2], detach_a[2], detach_d[2], attach[3];
  Enthalpies() {
    const double beta = 1.0;
    const double gP = 6.0, gX = 0.0, gE = 1.0;
    const double gCursor[4] = {-1.0, -1.0, -1.0, 1.5};  // A B C D
    const double fastest = (gCursor[cB] + gX) - (gCursor[cA] + gP);
    auto factor = [&](double g_left, double g_right) {
      const double f = std::exp(-(beta * (g_right - g_left - fastest)));
      if (f > 1.001) throw std::runtime_error("Setup error: Delta-G-fastest not actually fastest.");
      return std::min(1.0, f);
    };
    auto pair = [&](double* out, double g_left, double g_right) {
      const double r = factor(g_left, g_right);
      out[0] = r;
      out[1] = 1 - r;
    };
    for (int c = 0; c < 3; ++c) {
      pair(fwd[c], gCursor[c] + gP, gCursor[c + 1] + gX);
      pair(rev[c], gCursor[c + 1] + gX, gCursor[c] + gP);
    }
    pair(detach_a, gCursor[cA], gE);
    pair(detach_d, gCursor[cD], gE);
    const double to_a = factor(gE, gCursor[cA]), to_d = factor(gE, gCursor[cD]);
    if (to_a + to_d > 1.0)
      throw std::runtime_error("E->A+D rates too high to merge, given Delta-G-fastest.");
    attach[0] = to_a;
    attach[1] = to_d;
    attach[2] = 1.0 - to_a - to_d;
  }
};

void detachable(Machine& m) {
  static const Enthalpies g;
  const int fuel = m.read(P, 0);
  if (fuel == mP && bit(m.read(D, 1)) && bit(m.read(D, 2)) && coin(m)) {
    const int cursor = m.read(D, 0);
    for (int c = 0; c < 3; ++c) {
      if (cursor == kAdvance[c].from && m.pick(g.fwd[c], 2) == 0) {
        m.write(P, 0, mX);
        m.write(D, 0, kAdvance[c].leaves);
        m.write(D, 1, kAdvance[c].becomes);
        return;
      }
    }
    return;
  }
  if (fuel == mX && bit(m.read(D, -1)) && bit(m.read(D, -2))) {
    const int cursor = m.read(D, 0);
    for (int c = 0; c < 3; ++c) {
      if (cursor == kAdvance[c].becomes && m.pick(g.rev[c], 2) == 0) {
        m.write(P, 0, mP);
        m.write(D, 0, random_bit(m));
        m.write(D, -1, kAdvance[c].from);
        return;
      }
    }
    return;
  }
  if (fuel == mE && bit(m.read(D, 0)) && bit(m.read(D, +1)) && bit(m.read(D, -1)) && coin(m)) {
    const int outcome = m.pick(g.attach, 3);
    if (outcome == 2) return;
    m.write(P, 0, mS);
    m.write(D, 0, outcome == 0 ? cA : cD);
    return;
  }
  if (fuel == mS && bit(m.read(D, +1)) && bit(m.read(D, -1))) {
    const int cursor = m.read(D, 0);
    const double* leave = cursor == cA ? g.detach_a : (cursor == cD ? g.detach_d : nullptr);
    if (leave && m.pick(leave, 2) == 0) {
      m.write(P, 0, mE);
      m.write(D, 0, random_bit(m));
    }
  }
}
}  // namespace turing

// ---- ex5 family: alphabet (M S R T F), a four-instruction register machine. prob.scm:439-527 ----
namespace msrtf {
enum { opM = 0, opS = 1, opR = 2, opT = 3, opF = 4, opNone = -1 };

struct Registers {
  int budget = 4;        // Q: instructions still to fetch (> 0) or repeats done (< 0)
  int src = 0;           // Is
  int pc = 0;            // Ip
  int dst = 0;           // Id
  int last = opNone;     // Op
  bool seen_t = false, seen_r = false, seen_f = false;  // NT NR NF
};

void bump(Machine& m, int cell) { m.write(D, cell, (1 + m.read(D, cell)) % 5); }

void run(Machine& m, bool lone_r_executes) {
  Registers r;
  for (;;) {
    const int op = r.budget > 0 ? m.read(P, r.pc) : r.last;
    if (r.budget == 4) {  // first instruction: only S starts a program (prob.scm:445-450)
      if (op == opS) {
        r.budget = 3; r.pc += 1; r.last = op;
        r.seen_t = r.seen_r = r.seen_f = false;
        continue;
      }
      if (op == opR && lone_r_executes) bump(m, r.dst);
      return;
    }
    const bool final_slot = r.budget == 1 || r.budget == -3;
    const bool fetching = r.budget > 0;
    switch (op) {
      case opT: {
        const bool active = r.seen_t && r.seen_f;
        if (active) m.write(D, r.dst, m.read(P, r.src));
        if (final_slot) return;
        if (active) { r.src += 1; r.dst += 1; }
        r.seen_t = true;
        break;
      }
      case opR:
        if (r.seen_r) bump(m, r.dst);
        if (final_slot) return;
        r.seen_r = true;
        break;
      case opF:
        if (final_slot) return;
        r.seen_f = true;
        break;
      case opM:  // repeat the previous R/T three more times (prob.scm:479-481)
        if (r.last != opR && r.last != opT) return;
        r.budget = -1;
        continue;
      default:
        return;
    }
    if (fetching) r.pc += 1;
    r.budget -= 1;
    r.last = op;
  }
}
void strict(Machine& m) { run(m, false); }
void lenient(Machine& m) { run(m, true); }
}  // namespace msrtf

}  // namespace

void register_builtin_problems() {
  if (find_problem("ex1-radioactive-decay")) return;
  register_problem("__canary_problem_radioactive_decay", 2, decay);
  register_problem("ex1-radioactive-decay", 2, decay);
  register_problem("ex2-ferromagnetic-chain", 2, ferromagnet);
  register_problem("ex3-copolymerization", 4, copoly::plain);
  register_problem("ex3var1-copolymerization", 4, copoly::alternating);
  register_problem("ex3var2-copolymerization", 4, copoly::reversible);
  register_problem("ex4-chemical-turing", 9, turing::base);
  register_problem("ex4var1-chemical-turing", 9, turing::unsuppressed);
  register_problem("ex4var2-chemical-turing", 10, turing::detachable);
  register_problem("ex5-msrtf-machine", 5, msrtf::strict);
  register_problem("ex5var1-msrtf-machine", 5, msrtf::lenient);
}

}  // namespace tapes
