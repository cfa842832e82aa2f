"""API constants: label keys, env-var names, condition reasons, prefixes, phase tables.

Parity: /root/reference/pkg/apis/aitrainingjob/v1/constants.go:3-77 (label keys :3-11, env
names :13-21, reasons :23-39, ``aitj-`` prefixes :41-44, container error reasons :46-56,
ending phases :58-64, phase->reason map :65-77) and register.go:27-38 (group/version/kind).
Spellings that are visible in ``kubectl describe`` are kept verbatim (``Succeed``,
``TrainingJobSucceed``).
"""
from __future__ import annotations

# --- group / version / kind (register.go:27-38) ------------------------------------------
GROUP_NAME = "elasticdeeplearning.ai"
GROUP_VERSION = "v1"
API_VERSION = f"{GROUP_NAME}/{GROUP_VERSION}"
KIND = "AITrainingJob"
KIND_LIST = "AITrainingJobList"
KIND_PLURAL = "aitrainingjobs"
SHORT_NAME = "aitj"


def crd_name() -> str:
    return f"{KIND_PLURAL}.{GROUP_NAME}"


# --- labels (constants.go:3-11) -------------------------------------------------------------
CONTROLLER_NAME = "TrainingJobOperator"
LABEL_REPLICA_NAME = "TrainingJobReplicaName"
LABEL_REPLICA_INDEX = "TrainingJobReplicaIndex"
LABEL_JOB_NAME = "TrainingJobName"
LABEL_FRAMEWORK = "FrameworkType"
LABEL_GROUP_NAME = "GroupName"
LABEL_PRIORITY = "priority"
# extra labels written by the pod builder (pod.go:496-500)
LABEL_JOBNAME_COMPAT = "JobName"
LABEL_POD_ROLE = "PodRole"
LABEL_RESTART_COUNT = "RestartCount"
# new in this framework: rendezvous generation a replica was created for (elastic rescale)
LABEL_GENERATION = "RendezvousGeneration"

# --- env contract (constants.go:13-21) -----------------------------------------------------
ENV_REPLICA_NAME = "TRAININGJOB_REPLICA_NAME"
ENV_REPLICA_INDEX = "TRAININGJOB_REPLICA_INDEX"
ENV_REPLICA_RESTARTCOUNT = "TRAININGJOB_REPLICA_RESTARTCOUNT"
ENV_JOB_NAME = "TRAININGJOB_NAME"
ENV_JOB_NAMESPACE = "TRAININGJOB_NAMESPACE"
ENV_SERVICE = "TRAININGJOB_SERVICE"
ENV_PORTS = "TRAININGJOB_PORTS"

# --- reasons (constants.go:23-39) ------------------------------------------------------------
REASON_POD_TEMPLATE_RESTART_POLICY = "SettedPodTemplateRestartPolicy"
REASON_EXITED_WITH_CODE = "ExitedWithCode"

# --- phases (types.go:98-124) -----------------------------------------------------------------
PHASE_NONE = ""
PHASE_PENDING = "Pending"
PHASE_CREATING = "Creating"
PHASE_RUNNING = "Running"
PHASE_SUCCEEDED = "Succeed"  # sic
PHASE_FAILED = "Failed"
PHASE_TIMEOUT = "Timeout"
PHASE_RESTARTING = "Restarting"
PHASE_TERMINATING = "Terminating"
PHASE_PREEMPTED = "Preempted"
PHASE_NODE_FAIL = "NodeFail"

ALL_PHASES = (PHASE_NONE, PHASE_PENDING, PHASE_CREATING, PHASE_RUNNING, PHASE_SUCCEEDED, PHASE_FAILED,
              PHASE_TIMEOUT, PHASE_RESTARTING, PHASE_TERMINATING, PHASE_PREEMPTED, PHASE_NODE_FAIL)

TRAINING_JOB_REASON = {
    PHASE_NONE: "",
    PHASE_PENDING: "TrainingJobPending",
    PHASE_CREATING: "TrainingJobCreating",
    PHASE_RUNNING: "TrainingJobRunning",
    PHASE_SUCCEEDED: "TrainingJobSucceed",
    PHASE_FAILED: "TrainingJobFailed",
    PHASE_TIMEOUT: "TrainingJobTimeout",
    PHASE_RESTARTING: "TrainingJobRestarting",
    PHASE_TERMINATING: "TrainingJobTerminating",
    PHASE_PREEMPTED: "TrainingJobPreempted",
    PHASE_NODE_FAIL: "TrainingJobNodeFail",
}

# phases after which a job is finished (constants.go:58-64)
ENDING_PHASES = (PHASE_SUCCEEDED, PHASE_FAILED, PHASE_TIMEOUT, PHASE_PREEMPTED, PHASE_NODE_FAIL)
# phases in which the sync handler still reconciles (controller.go:298-304)
RECONCILABLE_PHASES = (PHASE_NONE, PHASE_PENDING, PHASE_CREATING, PHASE_RUNNING, PHASE_RESTARTING,
                       PHASE_TERMINATING)

# --- name prefixes (constants.go:41-44) ------------------------------------------------------
DEFAULT_CONTAINER_PREFIX = "aitj-"
DEFAULT_PORT_PREFIX = "aitj-"

# --- container start errors that count as "creating failed" (constants.go:46-56) ---------------
ERROR_CONTAINER_STATUS = (
    "CreateContainerConfigError",
    "CreateContainerError",
    "ImagePullBackOff",
    "ImageInspectError",
    "ErrImagePull",
    "ErrImageNeverPull",
    "RegistryUnavailable",
    "InvalidImageName",
)

# --- enums (types.go:64-72, replica.go:22-63) ---------------------------------------------------
CLEAN_POD_POLICY_ALL = "All"
CLEAN_POD_POLICY_NONE = "None"
CLEAN_POD_POLICIES = (CLEAN_POD_POLICY_ALL, CLEAN_POD_POLICY_NONE)

RESTART_POLICY_ALWAYS = "Always"
RESTART_POLICY_ON_FAILURE = "OnFailure"
RESTART_POLICY_ON_NODE_FAIL = "OnNodeFail"
RESTART_POLICY_NEVER = "Never"
RESTART_POLICY_EXIT_CODE = "ExitCode"
RESTART_POLICY_ON_NODE_FAIL_WITH_EXIT_CODE = "OnNodeFailWithExitCode"
RESTART_POLICIES = (RESTART_POLICY_ALWAYS, RESTART_POLICY_ON_FAILURE, RESTART_POLICY_ON_NODE_FAIL,
                    RESTART_POLICY_NEVER, RESTART_POLICY_EXIT_CODE, RESTART_POLICY_ON_NODE_FAIL_WITH_EXIT_CODE)

RESTART_SCOPE_ALL = "All"
RESTART_SCOPE_REPLICA = "Replica"
RESTART_SCOPE_POD = "Pod"
RESTART_SCOPES = (RESTART_SCOPE_ALL, RESTART_SCOPE_REPLICA, RESTART_SCOPE_POD)

ENDING_POLICY_ALL = "All"
ENDING_POLICY_RANK0 = "Rank0"
ENDING_POLICY_ANY = "Any"
ENDING_POLICY_NONE = "None"
ENDING_POLICIES = (ENDING_POLICY_ALL, ENDING_POLICY_RANK0, ENDING_POLICY_ANY, ENDING_POLICY_NONE)

EDL_POLICY_AUTO = "Auto"
EDL_POLICY_MANUAL = "Manual"
EDL_POLICY_NEVER = "Never"
EDL_POLICIES = (EDL_POLICY_AUTO, EDL_POLICY_MANUAL, EDL_POLICY_NEVER)

# --- pod phases / conditions of the core API subset we serve ------------------------------------
POD_PENDING = "Pending"
POD_RUNNING = "Running"
POD_SUCCEEDED = "Succeeded"
POD_FAILED = "Failed"
POD_UNKNOWN = "Unknown"

# --- annotations used by the single-box runtime (new; not in the reference) --------------------
ANN_GPUS = "aitj.b200/gpus"                      # GPUs bound to a pod ("0" or "2,3")
ANN_INJECT_FAULT = "aitj.b200/inject-fault"      # on a Node: mark the GPU unhealthy (fault injection)
ANN_RENDEZVOUS_PORT = "aitj.b200/master-port"    # on a pod: MASTER_PORT it was created with
ANN_SCALE_DOWN = "aitj.b200/scale-down"          # on a pod: draining because replicas shrank
ANN_TRACE = "aitj.b200/trace"                    # on a job: JSON timeline of lifecycle timestamps
