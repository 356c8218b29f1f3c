"""Scheduler class names (the harness subclasses them around oracle/sched_oracle.py)."""


class SchedulerMixin:
    order = 1
    init_noise_sigma = 1.0


class FlowMatchEulerDiscreteScheduler(SchedulerMixin):
    pass


class UniPCMultistepScheduler(SchedulerMixin):
    pass


class CogVideoXDDIMScheduler(SchedulerMixin):
    pass


class CogVideoXDPMScheduler(SchedulerMixin):
    pass
