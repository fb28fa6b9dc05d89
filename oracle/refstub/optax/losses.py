"""optax.losses restated (optax 0.2.4, _src/loss/_classification.py): kl_divergence, convex_kl_divergence"""
import jax.numpy as jnp


def kl_divergence(log_predictions, targets, axis=-1, where=None):
    loss = targets * (jnp.where(targets == 0, 0.0, jnp.log(jnp.where(targets == 0, 1.0, targets))) - log_predictions)
    return jnp.sum(loss, axis=axis)


def convex_kl_divergence(log_predictions, targets, axis=-1, where=None):
    """kl_divergence + sum(exp(log_predictions) - targets): jointly convex for unnormalised inputs"""
    return kl_divergence(log_predictions, targets, axis=axis) + jnp.sum(jnp.exp(log_predictions) - targets, axis=axis)


def safe_softmax_cross_entropy(logits, labels):
    import jax

    lp = jax.nn.log_softmax(logits, axis=-1)
    return -jnp.sum(jnp.where(labels == 0, 0.0, labels * lp), axis=-1)


def softmax_cross_entropy(logits, labels):
    import jax

    return -jnp.sum(labels * jax.nn.log_softmax(logits, axis=-1), axis=-1)


def l2_loss(predictions, targets=None):
    d = predictions if targets is None else predictions - targets
    return 0.5 * d * d
