// plugins/cartpole.cuh — the cart-pole of BASELINE config C4 as a PLUG-IN model: this text is not
// compiled into libaltro_b200.so; the library instantiates its kernel templates on it with NVRTC the
// first time a solver for ALTRO_B200_MODEL_CARTPOLE is created (csrc/modules.inl), and caches the
// module.  It is also the template for user models: write a struct with the model concept of
// csrc/device.cuh, hand its text to altro_b200_register_model, use the returned id with
// altro_b200_problem_set_model.
//
// Frictionless cart-pole, state (x, theta, xdot, thetadot), control = force on the cart, parameters
// P = (m_cart, m_pole, l, g); theta = 0 hangs down.  Definition owned by this repo;
// the CPU checker used by the tests restates the same equations independently.
struct Cartpole {  // definition owned by this repo
  static constexpr int n = 4, m = 1;
  static constexpr bool kDiscrete = false;
  static constexpr bool kStage3RepeatsStage2 = false;
  static __device__ __forceinline__ void eval(const double* P, const double* x, const double* u,
                                              double* xd) {
    const double mc = P[0], mp = P[1], l = P[2], g = P[3];
    const double thd = x[3];
    double s, c;
    sincos(x[1], &s, &c);
    const double den = mc + mp * s * s;
    const double F = u[0];
    xd[0] = x[2];
    xd[1] = thd;
    xd[2] = (F + mp * s * (l * thd * thd + g * c)) / den;
    xd[3] = (-F * c - mp * l * thd * thd * c * s - (mc + mp) * g * s) / (l * den);
  }
  static __device__ __forceinline__ void jac(const double* P, const double* x, const double* u,
                                             double* A, double* B) {
    const double mc = P[0], mp = P[1], l = P[2], g = P[3];
    const double thd = x[3];
    double s, c;
    sincos(x[1], &s, &c);
    const double den = mc + mp * s * s;
    const double F = u[0];
    const double numx = F + mp * s * (l * thd * thd + g * c);
    const double numt = -F * c - mp * l * thd * thd * c * s - (mc + mp) * g * s;
    const double dden = 2.0 * mp * s * c;
    const double dnumx = mp * c * (l * thd * thd + g * c) - mp * s * g * s;
    const double dnumt = F * s - mp * l * thd * thd * (c * c - s * s) - (mc + mp) * g * c;
    ALTRO_UNROLL
    for (int i = 0; i < 16; ++i) A[i] = 0.0;
    A[0 + 2 * 4] = 1.0;
    A[1 + 3 * 4] = 1.0;
    A[2 + 1 * 4] = (dnumx * den - numx * dden) / (den * den);
    A[2 + 3 * 4] = (2.0 * mp * s * l * thd) / den;
    A[3 + 1 * 4] = (dnumt * den - numt * dden) / (l * den * den);
    A[3 + 3 * 4] = (-2.0 * mp * l * thd * c * s) / (l * den);
    B[0] = 0.0;
    B[1] = 0.0;
    B[2] = 1.0 / den;
    B[3] = -c / (l * den);
  }
};
