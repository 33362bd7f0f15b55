"""oracle/tfemu is what lets the reference's own files run here (tests/golden/make_golden_reference.py); the reference
pins are only as good as its restatement of the TensorFlow primitives.  One test per restated rule, each against the
closed form TensorFlow documents (sources cited in oracle/tfemu/tensorflow/__init__.py).  CPU tier."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tf():
    # loaded under a private name: `import tensorflow` must keep failing in this process (bench.probe_reference)
    spec = importlib.util.spec_from_file_location("_tfemu_tensorflow_t", os.path.join(ROOT, "oracle", "tfemu", "tensorflow", "__init__.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_emulation_does_not_shadow_tensorflow(tf):
    import sys
    assert "tensorflow" not in sys.modules or not getattr(sys.modules["tensorflow"], "__version__", "").endswith("emu")


def test_numpy_operands_take_the_tensor_dtype(tf):
    t = tf.constant(np.ones((2, 3), np.float32))
    for expr in (t * np.full((2, 3), 1 / 3), np.full((2, 3), 1 / 3) * t, np.log(2.0) - t, 2.0 * t, t / np.float64(3.0),
                 (1 - np.zeros(3)) * t):
        assert isinstance(expr, tf.Tensor) and expr.dtype == tf.float32
    # the float64 operand is ROUNDED to float32 first (std * u in continuous_actors.py:296)
    u = np.array([1 / 3], np.float64)
    assert (tf.constant(np.array([3.0], np.float32)) * u).numpy()[0] == np.float32(3.0) * np.float32(1 / 3)
    with pytest.raises(TypeError):
        t * tf.constant(np.ones((2, 3), np.float64))                    # two tensors of different dtypes: TF raises
    assert tf.square(np.ones(2, np.float64)).dtype == tf.float64        # no tensor operand: NumPy's own dtype
    assert tf.math.log(2 * np.pi).dtype == tf.float32                   # Python float -> float32
    assert np.shape(t) == (2, 3) and isinstance(np.stack((t, t), -1), np.ndarray)


def test_concat_and_reduce_over_python_sequences(tf):
    a = np.ones((2, 2), np.float64)
    b = tf.constant(np.ones((2, 1), np.float32))
    assert tf.concat([a, b], -1).dtype == tf.float32                    # dtype of the first Tensor in the list
    q = (tf.constant([1.0, 5.0]), tf.constant([2.0, 3.0]))
    assert tf.reduce_min(q, axis=0).numpy().tolist() == [1.0, 3.0]


def test_tie_and_boundary_gradients(tf):
    x = tf.Variable(np.array([1.0, 2.0, 2.0], np.float32))
    y = tf.Variable(np.array([1.0, 3.0, 1.0], np.float32))
    with tf.GradientTape() as tape:
        m = tf.reduce_sum(tf.reduce_min((x, y), axis=0) * np.array([1.0, 10.0, 100.0]))
    gx, gy = tape.gradient(m, [x, y])
    assert gx.numpy().tolist() == [0.5, 10.0, 0.0] and gy.numpy().tolist() == [0.5, 0.0, 100.0]     # tie split equally
    with tf.GradientTape() as tape:
        m = tf.reduce_sum(tf.maximum(x, y))
    gx, gy = tape.gradient(m, [x, y])
    assert gx.numpy().tolist() == [1.0, 0.0, 1.0] and gy.numpy().tolist() == [0.0, 1.0, 0.0]        # tie -> first argument
    z = tf.Variable(np.array([-2.0, -1.0, 0.0, 1.0, 2.0], np.float32))
    with tf.GradientTape() as tape:
        m = tf.reduce_sum(tf.clip_by_value(z, -1.0, 1.0))
    assert tape.gradient(m, z).numpy().tolist() == [0.0, 1.0, 1.0, 1.0, 0.0]                        # closed interval
    with tf.GradientTape() as tape:
        m = tf.reduce_sum(tf.keras.activations.relu(z))
    assert tape.gradient(m, z).numpy().tolist() == [0.0, 0.0, 0.0, 1.0, 1.0]


def test_tape_structure_unconnected_sources_and_second_order(tf):
    a = tf.Variable(np.array([1.0, 2.0], np.float32))
    b = tf.Variable(3.0, dtype=tf.float32)
    c = tf.Variable(1.0, dtype=tf.float32)
    with tf.GradientTape() as tape:
        loss = tf.reduce_sum(tf.square(a)) * b
    (ga,), gb, gc = tape.gradient(loss, [[a], b, c])                     # nested sources, trpo.py:67
    assert ga.numpy().tolist() == [6.0, 12.0] and float(gb) == 5.0 and gc is None
    # Hessian-vector product through nested tapes (trpo.py:213-222): f = sum(a^3) -> H v = 6 a v
    v = np.array([0.5, -1.0], np.float32)
    with tf.GradientTape() as outer:
        with tf.GradientTape() as inner:
            f = tf.reduce_sum(a * a * a)
        g = inner.gradient(f, [a])
        gv = tf.reduce_sum(tf.concat([tf.reshape(x, [-1]) for x in g], -1) * v)
    (hv,) = outer.gradient(gv, [a])
    assert np.allclose(hv.numpy(), 6 * a.numpy() * v)
    assert float(tf.stop_gradient(tf.constant(2.0))) == 2.0


def test_dense_sequential_layout_and_weights(tf):
    nn = tf.keras.Sequential(name="t")
    nn.add(tf.keras.layers.Dense(4, kernel_initializer=tf.keras.initializers.Orthogonal(gain=np.sqrt(2)),
                                 activation=tf.keras.activations.tanh, input_shape=(3,)))
    nn.add(tf.keras.layers.Dense(2, kernel_initializer="glorot_uniform"))
    tv = nn.trainable_variables                                          # complete right after construction (nn_utils.py)
    assert [tuple(v.shape) for v in tv] == [(3, 4), (4,), (4, 2), (2,)]
    k0 = nn.get_weights()[0]
    assert np.allclose(k0 @ k0.T, 2 * np.eye(3), atol=1e-5)              # orthogonal rows, gain^2
    w = [np.arange(12, dtype=np.float32).reshape(3, 4) / 10, np.ones(4, np.float32), np.ones((4, 2), np.float32),
         np.zeros(2, np.float32)]
    nn.set_weights(w)
    x = np.array([[1.0, 0.0, -1.0]], np.float32)
    want = np.tanh(x @ w[0] + w[1]) @ w[2] + w[3]
    assert np.allclose(nn(x).numpy(), want, atol=1e-6)
    with pytest.raises(ValueError):
        nn.set_weights(w[:3])


def test_keras_adam_first_steps_closed_form(tf):
    """One and two dense steps against the formula of keras/optimizers/adam.py::update_step evaluated in float64."""
    v = tf.Variable(np.array([1.0, -2.0], np.float32))
    opt = tf.keras.optimizers.Adam(learning_rate=1e-2)
    g1, g2 = np.array([0.3, -4.0]), np.array([-0.1, 1.0])
    opt.apply_gradients(zip([tf.constant(g1.astype(np.float32))], [v]))
    m, s = 0.1 * g1, 0.001 * g1 ** 2
    th = np.array([1.0, -2.0]) - 1e-2 * np.sqrt(1 - 0.999) / (1 - 0.9) * m / (np.sqrt(s) + 1e-7)
    assert np.allclose(v.numpy(), th, rtol=2e-7, atol=0)
    opt.apply_gradients(zip([tf.constant(g2.astype(np.float32))], [v]))
    m, s = m + (g2 - m) * 0.1, s + (g2 ** 2 - s) * 0.001
    th = th - 1e-2 * np.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2) * m / (np.sqrt(s) + 1e-7)
    assert np.allclose(v.numpy(), th, rtol=3e-7, atol=0) and opt.iterations == 2
    # epsilon is NOT bias-corrected: for |g| << eps the step is lr_t * m / eps, not lr * sign(g)
    w = tf.Variable(np.array([0.0], np.float32))
    o2 = tf.keras.optimizers.Adam(learning_rate=1.0)
    o2.apply_gradients(zip([tf.constant(np.array([1e-12], np.float32))], [w]))
    assert abs(float(w.numpy()[0]) + np.sqrt(0.001) / 0.1 * 0.1 * 1e-12 / (np.sqrt(0.001) * 1e-12 + 1e-7)) < 1e-9


def test_clip_by_global_norm_and_norms(tf):
    gs = [tf.constant(np.array([3.0, 0.0], np.float32)), None, tf.constant(np.array([[4.0]], np.float32))]
    out, gn = tf.clip_by_global_norm(gs, 1.0)
    assert float(gn) == 5.0 and out[1] is None and np.allclose(out[0].numpy(), [0.6, 0.0]) and np.allclose(out[2].numpy(), [[0.8]])
    out, _ = tf.clip_by_global_norm(gs, 10.0)                             # below the threshold: unchanged
    assert np.allclose(out[0].numpy(), [3.0, 0.0])
    assert float(tf.linalg.global_norm([gs[0], gs[2]])) == 5.0
    assert np.allclose(tf.math.reduce_euclidean_norm(np.array([[3.0, 4.0], [6.0, 8.0]], np.float32), axis=1).numpy(), [5, 10])
    assert float(tf.norm(gs[0])) == 3.0


def test_variable_assign_squeeze_split_shapes(tf):
    v = tf.Variable(np.log(0.1), dtype=tf.float32)
    v.assign(np.maximum(v.numpy(), 1e-5))
    assert v.dtype == tf.float32 and float(v) == np.float32(1e-5)
    t = tf.constant(np.arange(12, dtype=np.float32).reshape(2, 6))
    a, b = tf.split(t, num_or_size_splits=2, axis=-1)
    assert tuple(a.shape) == (2, 3) and b.numpy()[0, 0] == 3.0
    assert tuple(tf.squeeze(tf.expand_dims(a, axis=-1), axis=-1).shape) == (2, 3)
    with pytest.raises(ValueError):
        tf.squeeze(a, axis=0)
    assert tf.shape(t).numpy().tolist() == [2, 6] and int(tf.size(t).numpy()) == 12
    assert len(t.shape) == 2 and t[:, :-1].shape == (2, 5) and t[:, -1].shape == (2,)
    x = np.array([-30.0, -1.0, 0.0, 2.0, 40.0], np.float32)
    assert np.allclose(tf.nn.softplus(x).numpy(), np.logaddexp(0, x.astype(np.float64)), rtol=1e-6)
