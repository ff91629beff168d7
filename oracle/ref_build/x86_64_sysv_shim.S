/*
 * x86_64_sysv_shim.S - TEST INFRASTRUCTURE ONLY (oracle build, never shipped).
 *
 * The reference (ambonvik/cimba) keeps its coroutine context switch and its
 * CPU-entropy probes in two NASM files:
 *   src/port/x86-64/linux/cmi_coroutine_context.asm:90-148
 *   src/port/x86-64/linux/cmi_random_hwseed.asm:28-66
 * This image has no `nasm`, so the oracle build (oracle/Makefile) assembles
 * this GNU-as file instead.  It provides the same six symbols with the same
 * calling contract for the -DNMXCSR frame layout that
 * src/port/x86-64/linux/cmi_coroutine_context.c:118-184 prepares
 * (8 slots: return address, rflags, rbp, rbx, r12, r13, r14, r15).
 *
 * Nothing in the B200 product path uses this file: the device engine has no
 * switchable stacks (processes are resume-point state machines).
 */
        .text

/* void *cmi_coroutine_context_switch(void **old_sp, void **new_sp, void *msg)
 *   rdi = where to park the outgoing stack pointer
 *   rsi = where to fetch the incoming stack pointer
 *   rdx = message handed to the incoming context as its return value
 */
        .globl  cmi_coroutine_context_switch
        .type   cmi_coroutine_context_switch, @function
cmi_coroutine_context_switch:
        pushfq
        pushq   %rbp
        pushq   %rbx
        pushq   %r12
        pushq   %r13
        pushq   %r14
        pushq   %r15
        movq    %rsp, (%rdi)
        movq    (%rsi), %rsp
        popq    %r15
        popq    %r14
        popq    %r13
        popq    %r12
        popq    %rbx
        popq    %rbp
        popfq
        movq    %rdx, %rax
        popq    %r9
        jmpq    *%r9
        .size   cmi_coroutine_context_switch, .-cmi_coroutine_context_switch

/* First activation of a coroutine lands here (address pre-loaded as the
 * "return address" of a fresh frame).  r12 = body, r13 = coroutine,
 * r14 = context argument, r15 = exit handler that receives the body's
 * return value if the body ever returns.
 */
        .globl  cmi_coroutine_trampoline
        .type   cmi_coroutine_trampoline, @function
cmi_coroutine_trampoline:
        movq    %r13, %rdi
        movq    %r14, %rsi
        xorq    %rax, %rax
        callq   *%r12
        pushq   %rdi
        movq    %rax, %rdi
        jmpq    *%r15
        .size   cmi_coroutine_trampoline, .-cmi_coroutine_trampoline

/* int cmi_cpu_has_rdseed(void): CPUID.(EAX=7,ECX=0):EBX bit 18 */
        .globl  cmi_cpu_has_rdseed
        .type   cmi_cpu_has_rdseed, @function
cmi_cpu_has_rdseed:
        pushq   %rbx
        movl    $7, %eax
        xorl    %ecx, %ecx
        cpuid
        movl    %ebx, %eax
        shrl    $18, %eax
        andl    $1, %eax
        popq    %rbx
        ret
        .size   cmi_cpu_has_rdseed, .-cmi_cpu_has_rdseed

/* int cmi_cpu_has_rdrand(void): CPUID.(EAX=1):ECX bit 30 */
        .globl  cmi_cpu_has_rdrand
        .type   cmi_cpu_has_rdrand, @function
cmi_cpu_has_rdrand:
        pushq   %rbx
        movl    $1, %eax
        cpuid
        movl    %ecx, %eax
        shrl    $30, %eax
        andl    $1, %eax
        popq    %rbx
        ret
        .size   cmi_cpu_has_rdrand, .-cmi_cpu_has_rdrand

/* uint64_t cmi_rdseed(void): spin (with pause) until the entropy pool delivers */
        .globl  cmi_rdseed
        .type   cmi_rdseed, @function
cmi_rdseed:
1:      rdseed  %rax
        jc      2f
        pause
        jmp     1b
2:      ret
        .size   cmi_rdseed, .-cmi_rdseed

/* uint64_t cmi_rdrand(void) */
        .globl  cmi_rdrand
        .type   cmi_rdrand, @function
cmi_rdrand:
1:      rdrand  %rax
        jnc     1b
        ret
        .size   cmi_rdrand, .-cmi_rdrand

        .section .note.GNU-stack,"",@progbits
